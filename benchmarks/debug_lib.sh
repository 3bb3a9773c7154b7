# Builds build/libamf_dbg.so = the whole library with -DAMF_BOUNDS_CHECK (device-side bounds checks,
# csrc/common.cuh) without touching the in-tree release build; run anything on it with
#   AMF_B200_LIB=build/libamf_dbg.so python -m pytest tests -m gpu
set -e
cd "$(dirname "$0")/.."
C=active_matrix_factorization_b200/csrc
mkdir -p build/dbg
for f in $C/*.cu; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    --expt-relaxed-constexpr -DAMF_BOUNDS_CHECK -c $f -o build/dbg/$b.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/libamf_dbg.so build/dbg/*.o -cudart static
echo build/libamf_dbg.so
