#!/bin/bash
# build a variant of the library for A/B runs: tools/build_variant.sh NAME [-DMACRO=..]...
# -> s3prl_b200/_lib/ab_NAME.so (objects under .ab/NAME)
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OBJ=$ROOT/.ab/$NAME; mkdir -p $OBJ
cd $ROOT/s3prl_b200/csrc
ls *.cu | xargs -P 8 -I{} nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c {} -o $OBJ/{}.o
nvcc -shared -o $ROOT/s3prl_b200/_lib/ab_$NAME.so $OBJ/*.o -gencode arch=compute_100a,code=sm_100a
echo built $ROOT/s3prl_b200/_lib/ab_$NAME.so
