#!/bin/bash
# A/B kernel variants: recompiles the tick kernels + C-ABI host code with extra -D flags into
# dnn-mppi-mpc_b200/mppi_b200/libmppi_b200_<name>.so (the other objects come from the default build).
# Usage: profiles/scripts/build_variant.sh <name> "<-D flags>"; run with MPPI_B200_LIB=<that .so>
set -e
cd "$(dirname "$0")/../../dnn-mppi-mpc_b200/csrc"
name=$1; flags=$2
mkdir -p build_$name
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ $flags"
$NV -c mppi_kernels.cu -o build_$name/mppi_kernels.o &
$NV -c mppi_api.cu -o build_$name/mppi_api.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../mppi_b200/libmppi_b200_$name.so build_$name/mppi_kernels.o build_$name/mppi_api.o \
    build/mppi_mlp.o build/mppi_topn.o build/mppi_spline.o build/mppi_probe.o -ldl
echo built $name
