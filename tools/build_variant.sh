#!/bin/bash
# A/B builds of the library: tools/build_variant.sh <name> <extra nvcc flags...>  ->  schnetpack_b200/csrc/libspk_b200_<name>.so
# (select at run time with SPK_B200_LIB=<path>)
set -e
name=$1; shift
cd "$(dirname "$0")/../schnetpack_b200/csrc"
mkdir -p build_$name
pids=()
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 550 "$@" -DSPK_TU=${f%.cu} -c $f -o build_$name/${f%.cu}.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o libspk_b200_$name.so build_$name/*.o -lcudart
echo built $(pwd)/libspk_b200_$name.so
