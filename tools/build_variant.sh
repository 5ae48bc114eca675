#!/bin/bash
# A/B builds of the same sources: tools/build_variant.sh NAME "-DMACRO=value ..."  ->  build_variants/libNAME.so
# (load it with LMPC_GPU_LIB=build_variants/libNAME.so; build_variants/ is git-ignored but travels with gpurun)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/build_variants/$1"
make -s -C "$ROOT/linearmpc.jl_b200/csrc" -j8 OBJ="$ROOT/build_variants/$1" OUT="$ROOT/build_variants/lib$1.so" EXTRA="$2"
