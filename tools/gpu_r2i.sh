#!/bin/bash
# round-2 pass i: clean A/B of the operand schemes on the current build (fused LN compiled out), lanes 1 vs 2
mkdir -p gpurun_out
for sch in bf16x3 f16q8; do
  S3B_GEMM_SCHEME=$sch timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_c2_$sch.json 2> gpurun_out/r2i_c2_$sch.err
  S3B_GEMM_SCHEME=$sch timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_c3_$sch.json 2> gpurun_out/r2i_c3_$sch.err
  S3B_GEMM_SCHEME=$sch timeout 300 python bench.py --steps 20 --warmup 3 --emulate-world 8 --no-cpu-baseline > gpurun_out/r2i_shard8_$sch.json 2> gpurun_out/r2i_shard8_$sch.err
done
S3B_GEMM_SCHEME=bf16x3 timeout 300 python bench.py --steps 20 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/r2i_c2_bf16x3_l1.json 2> gpurun_out/r2i_c2_bf16x3_l1.err
