#!/bin/bash
# A/B builds of the library on ONE box (GEMM time varies several % box to box with power/clocks):
#   tools/ab_step.sh "32 4" libA.so libB.so ...   ("-" = the in-tree build)
# alternates the libraries twice per batch size; prints ms/step of tools/profile_step.py
BATCHES=$1; shift
for b in $BATCHES; do
  for rep in 1 2; do
    for v in "$@"; do
      if [ "$v" = "-" ]; then unset S3B_LIB_PATH; else export S3B_LIB_PATH=$PWD/$v; fi
      echo -n "B=$b $v: "; timeout 200 python tools/profile_step.py --batch $b --steps 20 --warmup 5 2>&1 | grep "ms/step" | head -1
    done
  done
done
