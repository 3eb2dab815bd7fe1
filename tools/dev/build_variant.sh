#!/bin/bash
# Dev: builds libssim_cuda.so with extra compiler flags into build/var_<name>/ (for tools/dev/variant_ab.py) and prints the
# instruction counts of the two hot loops.   tools/dev/build_variant.sh NAME "-DSOME_MACRO=1 ..."
set -e
cd "$(dirname "$0")/../.."
N=$1; FL=$2
D=build/var_$N
mkdir -p $D
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -Iinclude -Issim_b200/csrc $FL"
nvcc $NVF -c ssim_b200/csrc/ssim_kernels.cu -o $D/ssim_kernels.o &
nvcc $NVF -c ssim_b200/csrc/ssim_cuda.cu -o $D/ssim_cuda.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o $D/libssim_cuda.so $D/ssim_kernels.o $D/ssim_cuda.o -cudart static -ldl -lpthread
python tools/sass_summary.py $D/libssim_cuda.so | grep -A4 "fused_kernel<1, false>"
