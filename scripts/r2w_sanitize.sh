# compute-sanitizer memcheck over what the second half of round 2 added: sphere / cube trees, overlapped frames, overlapped path frames
# (racecheck inspects shared-memory hazards only; these kernels use none beyond what r2s_sanitize.sh already covers)
mkdir -p gpurun_out
SEL='tie_rules or frame_with_many_primitives or overlapped_frames_identical and 4 or overlapped_path_frames_identical and None'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/r2w_memcheck.log 2>&1
echo "memcheck rc=$?" | tee gpurun_out/r2w_summary.log
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2w_memcheck.log | tail -3 | tee -a gpurun_out/r2w_summary.log
