#!/bin/bash
# build_variant.sh NAME "-DPT_X=.. -DPT_Y=.." : compile pt_wavefront.cu with extra defines into expt/libptb200_NAME.so (A/B timing, tools/abtest.py)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p expt build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Ismall-pathtracer_b200/csrc -Xptxas -v $@ \
     -c small-pathtracer_b200/csrc/pt_wavefront.cu -o build/pt_wavefront_$NAME.o 2> build/v_$NAME.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static build/pt_validate.o build/pt_wavefront_$NAME.o build/pt_api.o build/pt_jit.o -ldl -o expt/libptb200_$NAME.so
grep -A2 "k_bounceILi[01]ELb0" build/v_$NAME.log | grep "spill\|Used" | tr '\n' ' '; echo
