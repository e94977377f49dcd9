# every bench.py workload on one B200, plus the full GPU test suite and the reference arm (round-2 final numbers)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2m_gputests.log
for w in soup1m niels360 niels1080 heightfield10m heightfield10m_b4 soup1m_far niels8k16 soup8k16; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 2>gpurun_out/r2m_$w.err > gpurun_out/r2m_bench_$w.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2m_bench_$w.json"))
    r = d["roofline"]
    print("$w", round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 3), "ms  median", round(d["ms_per_step_median"], 3), " e2e", round(d["e2e"]["value"], 1),
          " roof", r["bound"], round(r["frac"], 3), r["traffic"], " cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("$w failed", e)
PY
done
timeout 900 python bench.py --impl reference --steps 10 --warmup 1 2>gpurun_out/r2m_reference.err | tee gpurun_out/r2m_bench_reference.json | cut -c1-400
timeout 600 python bench.py --impl reference --workload niels360 --steps 10 --warmup 1 2>>gpurun_out/r2m_reference.err | tee gpurun_out/r2m_bench_reference_niels360.json | cut -c1-300
