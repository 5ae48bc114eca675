#!/bin/bash
# build_variant.sh NAME "EXTRA NVCC FLAGS": a second build of the library under build_variants/NAME.so
# (select it at run time with LMPC_GPU_LIB=build_variants/NAME.so) for A/B kernel measurements.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/linearmpc.jl_b200/csrc
OUT=$ROOT/build_variants
mkdir -p $OUT/obj_$1
FLAGS="-gencode arch=compute_100a,code=sm_100a $2 -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I$ROOT/include -I$SRC"
for f in api ldp_thread ldp_box ldp_warp ldp_cta bst_eval; do
  nvcc $FLAGS -c $SRC/$f.cu -o $OUT/obj_$1/$f.o &
done
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -I$ROOT/include -I$SRC -c $SRC/setup.cpp -o $OUT/obj_$1/setup.o
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -I$ROOT/include -I$SRC -c $SRC/condense.cpp -o $OUT/obj_$1/condense.o
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/$1.so $OUT/obj_$1/*.o -cudart static
rm -rf $OUT/obj_$1
echo built $OUT/$1.so
