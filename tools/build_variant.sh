#!/bin/bash
# Build an experimental variant of libofc.so (flow_kernels.cu compiled with -DOFC_EXP=<n>) into build_variants/libofc_exp<n>.so;
# tools/ab_flow.sh swaps it in on the GPU box ("name@exp<n>" specs).  Other sources are compiled once and cached as objects.
# (Wrap the code under test in `#if OFC_EXP == <n>` in flow_kernels.cu; nothing in the tree uses the macro between experiments.)
set -e
N=$1
cd "$(dirname "$0")/.."
mkdir -p build_variants/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
for f in opticalflowclustering_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  [ "$b" = flow_kernels ] && continue
  o=build_variants/obj/$b.o
  if [ ! -f $o ] || [ $f -nt $o ] || [ -n "$(find opticalflowclustering_b200/csrc -name '*.cuh' -newer $o)" ]; then
    nvcc $FLAGS -c -o $o $f &
  fi
done
nvcc $FLAGS -DOFC_EXP=$N -c -o build_variants/obj/flow_kernels_exp$N.o opticalflowclustering_b200/csrc/flow_kernels.cu &
wait
nvcc -shared -o build_variants/libofc_exp$N.so build_variants/obj/flow_kernels_exp$N.o $(ls build_variants/obj/*.o | grep -v flow_kernels_exp)
echo built build_variants/libofc_exp$N.so
