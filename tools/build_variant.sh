#!/bin/bash
# tools/build_variant.sh NAME FILE.cu "EXTRA NVCC FLAGS" -- A/B build: recompile one source with extra flags and link
# yagi_b200/lib/libyagi_b200_NAME.so from it plus the stock objects; select it at run time with YG_LIB=NAME.
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; extra=$3
python -c "from yagi_b200 import build; build.build()" > /dev/null
obj=yagi_b200/build/ab_${name}_${src%.cu}.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas=-v $extra -c yagi_b200/csrc/$src -o $obj 2> yagi_b200/build/ab_${name}.ptxas.log
objs=$(ls yagi_b200/build/*.o | grep -v "/ab_" | grep -v "/${src%.cu}.o")
nvcc -shared -o yagi_b200/lib/libyagi_b200_${name}.so $objs $obj -gencode arch=compute_100a,code=sm_100a -cudart static
echo built yagi_b200/lib/libyagi_b200_${name}.so
