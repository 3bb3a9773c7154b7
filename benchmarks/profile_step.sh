# Profiling recipe of the C5 step (B200_PROFILING.md): plain run first, then the ncu launch list of
# the same command, then one --set full capture of the two dominant kernels.
# usage: bash benchmarks/profile_step.sh <tag>      (outputs gpurun_out/<tag>_*)
set -x
T=${1:-r03}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-configs"
$B > gpurun_out/${T}_plain.json 2> gpurun_out/${T}_plain.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $B > gpurun_out/${T}_ncu_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'tiled_side_kernel|pool_pred_kernel' --launch-skip 9 -c 3 -o gpurun_out/${T}_main -f $B > gpurun_out/${T}_ncu_f.log 2>&1
ncu -i gpurun_out/${T}_main.ncu-rep --page raw --csv > gpurun_out/${T}_main_ncu_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_main.ncu-rep --page details > gpurun_out/${T}_details.txt 2>/dev/null
ls -la gpurun_out/ | tail -8
