# fresh --set full capture of the shading launches (after the light cache) and of the two traversal launches of the headline frame
mkdir -p gpurun_out
prof() { # workload skip count regex tag
  ncu --set full --clock-control none --import-source on -k regex:$4 --launch-skip $2 -c $3 -f -o /tmp/r2z_$5 python scripts/profile_frame.py --workload $1 --frames 3 > gpurun_out/r2z_$5.log 2>&1
  ncu -i /tmp/r2z_$5.ncu-rep --page raw --csv > gpurun_out/r2z_$5.csv 2>>gpurun_out/r2z_$5.log
  tail -2 gpurun_out/r2z_$5.log
}
prof soup1m 4 2 "k_shade|k_shadowgen" soup1m_shade
prof soup1m 4 2 k_trace soup1m
