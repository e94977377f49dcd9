# compute-sanitizer over the kernels added in round 2 (small frames): memcheck, then racecheck
mkdir -p gpurun_out
SEL='shadow_order_modes_identical and niels or path_frame_equals_the_oracle_statement_on_nielsscene and 1-4 or device_builder and (niels or two_triangles or clustered) or frame_graph_replay or test_cuda_equals_golden and release and 1-'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/r2s_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r2s_summary.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/r2s_$tool.log | tail -3 | tee -a gpurun_out/r2s_summary.log
done
