#!/bin/bash
# A/B builds of the product library with extra -D flags: tools/build_variant.sh NAME "-DSDG_X=1 ..." -> subrosadg_b200/libNAME.so
# (select with SDG_LIB=subrosadg_b200/libNAME.so; see tools/gpu_ab*.sh)
set -e
cd "$(dirname "$0")/../subrosadg_b200/csrc"
NAME=$1; shift
mkdir -p _build
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fopenmp,-O3,-Wall,-Wno-unknown-pragmas -Xptxas -v"
for tu in sdg_api euler_launch ns_launch nsl_launch; do
  ($NV $@ -c $tu.cu -o _build/${tu}_$NAME.o 2> _build/${tu}_$NAME.ptxas.log || (tail -30 _build/${tu}_$NAME.ptxas.log; exit 1)) &
done
wait
[ -f _build/mixed_path.o ] || $NV -c mixed_path.cu -o _build/mixed_path.o 2> _build/mixed_path.ptxas.log
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fopenmp -o ../lib$NAME.so _build/sdg_api_$NAME.o _build/euler_launch_$NAME.o _build/ns_launch_$NAME.o _build/nsl_launch_$NAME.o _build/mixed_path.o _build/physics_debug.o _build/view_variable.o -lgomp
grep -A3 "nsStageKernelILi3ELi4ELi4ELb1ELi1E" _build/ns_launch_$NAME.ptxas.log | grep "spill\|Used" | head -2
