# round-2 final artefacts on one B200: GPU tests, every bench.py workload, the reference arm, smoke(), the launch list of the bench
# command and a fresh --set full capture of the path-tracing traversal launches (camera of the final heightfield10m_b4 workload)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2z_gputests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2z_smoke.log
for w in soup1m niels360 niels1080 heightfield10m heightfield10m_b4 soup1m_far niels8k16 soup8k16; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 2>gpurun_out/r2z_$w.err > gpurun_out/r2z_bench_$w.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2z_bench_$w.json"))
    r = d["roofline"]
    print("$w", round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 3), "ms  median", round(d["ms_per_step_median"], 3), " e2e", round(d["e2e"]["value"], 1),
          " roof", r["bound"], round(r["frac"], 3), r["traffic"], " cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("$w failed", e)
PY
done
timeout 900 python bench.py --impl reference --steps 10 --warmup 1 2>gpurun_out/r2z_reference.err | tee gpurun_out/r2z_bench_reference_soup1m.json | cut -c1-300
timeout 600 python bench.py --impl reference --workload niels360 --steps 10 --warmup 1 2>>gpurun_out/r2z_reference.err | tee gpurun_out/r2z_bench_reference_niels360.json | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:k_trace --launch-skip 10 -c 10 -f -o /tmp/r2z_b4 python scripts/profile_frame.py --workload heightfield10m_b4 --frames 2 > gpurun_out/r2z_b4.log 2>&1
ncu -i /tmp/r2z_b4.ncu-rep --page raw --csv > gpurun_out/r2z_heightfield10m_b4.csv 2>>gpurun_out/r2z_b4.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_launches_bench.log 2>&1
ls -la gpurun_out/r2z_* | head -40
