mkdir -p gpurun_out
python -m pytest tests/test_gpu_path.py -m gpu -x -q -k 10m 2>&1 | tail -5 | tee gpurun_out/r2g_pathtest.log
prof() { # workload skip count regex tag
  ncu --set full --clock-control none --import-source on -k regex:$4 --launch-skip $2 -c $3 -f -o /tmp/r2g_$5 python scripts/profile_frame.py --workload $1 --frames 2 > gpurun_out/r2g_$5.log 2>&1
  ncu -i /tmp/r2g_$5.ncu-rep --page raw --csv > gpurun_out/r2g_$5.csv 2>>gpurun_out/r2g_$5.log
  tail -2 gpurun_out/r2g_$5.log
}
prof soup1m 2 2 k_trace soup1m
prof soup1m 2 2 "k_shade|k_shadowgen" soup1m_shade
prof heightfield10m_b4 10 10 k_trace heightfield10m_b4
prof heightfield10m 2 2 k_trace heightfield10m
prof niels1080 2 2 k_trace niels1080
prof soup8k16 2 2 k_trace soup8k16
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_launches_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches_b4.csv python scripts/profile_frame.py --workload heightfield10m_b4 --frames 3 > /dev/null 2>&1
ls -la gpurun_out/r2g_* | head -30
