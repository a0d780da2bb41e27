#!/bin/bash
# usage: scripts/build_variant.sh NAME -DNB_X=..   -> nimble_aligner_b200/libnimble_b200_NAME.so (kernel tuning experiments;
# select with NIMBLE_B200_SO=<path> when importing nimble_aligner_b200)
set -e
cd "$(dirname "$0")/../nimble_aligner_b200"
name=$1; shift
python build.py > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda -Xcompiler -fPIC,-pthread -cudart shared "$@" -c csrc/kernels.cu -o /tmp/kernels_$name.o
objs=$(ls csrc/*.o | grep -v kernels.cu.o)
nvcc -shared -cudart shared -o libnimble_b200_$name.so $objs /tmp/kernels_$name.o -lz -lpthread -ldl
echo libnimble_b200_$name.so
