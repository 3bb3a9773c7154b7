# Builds build/libamf_<tag>.so = the library with ONE translation unit recompiled with extra nvcc
# flags (kernel-variant timing; select it with AMF_B200_LIB=build/libamf_<tag>.so).
# usage: variant_lib.sh <tag> <file.cu> <extra nvcc flags...>
set -e
cd "$(dirname "$0")/.."
tag=$1; src=$2; shift 2
C=active_matrix_factorization_b200/csrc
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr "$@" -c $C/$src -o build/${src%.cu}_$tag.o
objs=""
for f in $C/*.cu; do b=$(basename $f .cu); if [ "$b.cu" = "$src" ]; then objs="$objs build/${b}_$tag.o"; else objs="$objs $C/$b.o"; fi; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/libamf_$tag.so $objs -cudart static
echo build/libamf_$tag.so
