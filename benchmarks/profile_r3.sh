# ncu captures of the two bundled-runs kernels on the C5 shapes (--set full); usage: profile_r3.sh TAG [pool|tiled|both]
set -x
T=${1:-r3}
W=${2:-both}
if [ "$W" != tiled ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_pred_kernel' --launch-skip 3 -c 1 -o gpurun_out/${T}_pool -f python benchmarks/pool_variants.py --child 224 > gpurun_out/${T}_pool_ncu.log 2>&1
fi
if [ "$W" != pool ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tiled_side_kernel' --launch-skip 6 -c 2 -o gpurun_out/${T}_tiled -f python benchmarks/tiled_variants.py --child tiled > gpurun_out/${T}_tiled_ncu.log 2>&1
fi
ls -la gpurun_out | tail -5
