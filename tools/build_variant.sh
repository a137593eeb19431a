#!/bin/bash
# Build a variant of the product library with extra nvcc flags for mpnn_tc.cu (kernel experiments):
#   tools/build_variant.sh NAME -DECO_X=1 ...   ->  eco-dqn_b200/lib/variants/libNAME.so  (select with ECO_DQN_B200_LIB)
set -e
cd "$(dirname "$0")/../eco-dqn_b200"
name=$1; shift
mkdir -p lib/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I ../include -I csrc "$@" -Xptxas -v -c csrc/mpnn_tc.cu -o lib/variants/mpnn_tc_$name.o 2>&1 | grep -A2 "ILb0ELb0ELb0" | grep -E "registers|spill" || true
objs=$(ls lib/*.o | grep -v -e mpnn_tc.o -e tc_probe.o)
/usr/local/cuda/bin/nvcc -shared -o lib/variants/lib$name.so $objs lib/variants/mpnn_tc_$name.o -gencode arch=compute_100a,code=sm_100a
rm lib/variants/mpnn_tc_$name.o
echo built lib/variants/lib$name.so
