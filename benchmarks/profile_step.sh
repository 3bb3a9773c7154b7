set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$B > gpurun_out/r01i_plain.json 2> gpurun_out/r01i_plain.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01i_launches.csv $B > gpurun_out/r01i_ncu_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'tiled_side_kernel|pool_pred_kernel' --launch-skip 9 -c 3 -o gpurun_out/prof_r01i -f $B > gpurun_out/r01i_ncu_f.log 2>&1
ncu -i gpurun_out/prof_r01i.ncu-rep --page raw --csv > gpurun_out/r01i_ncu_full_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r01i.ncu-rep --page details > gpurun_out/r01i_details.txt 2>/dev/null
ls -la gpurun_out/ | tail -8
