#!/bin/bash
# round-2 pass f: clean A/B of the operand schemes and of the fused LayerNorm (fc2 only, two rows at a time)
mkdir -p gpurun_out
S3B_GEMM_SCHEME=f16q8 timeout 500 python tools/gemm_tile_sweep.py > gpurun_out/r2f_gemm_tile_sweep_f16q8.txt 2>&1
for sch in f16q8 bf16x3; do
  for fuse in 0 2; do
    if [ $fuse = 2 ]; then unset S3B_FUSE_LN; else export S3B_FUSE_LN=0; fi
    S3B_GEMM_SCHEME=$sch timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_c2_${sch}_fuse$fuse.json 2> gpurun_out/r2f_c2_${sch}_fuse$fuse.err
  done
done
unset S3B_FUSE_LN
S3B_GEMM_SCHEME=f16q8 timeout 300 python bench.py --steps 20 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/r2f_c2_f16q8_l1.json 2> gpurun_out/r2f_c2_f16q8_l1.err
S3B_GEMM_SCHEME=f16q8 timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_c3_f16q8.json 2> gpurun_out/r2f_c3_f16q8.err
S3B_GEMM_SCHEME=f16q8 timeout 300 python bench.py --steps 20 --warmup 3 --emulate-world 8 --no-cpu-baseline > gpurun_out/r2f_shard8_f16q8.json 2> gpurun_out/r2f_shard8_f16q8.err
S3B_GEMM_SCHEME=bf16x3 timeout 300 python bench.py --steps 20 --warmup 3 --emulate-world 8 --no-cpu-baseline > gpurun_out/r2f_shard8_bf16x3.json 2> gpurun_out/r2f_shard8_bf16x3.err
S3B_GEMM_SCHEME=f16q8 timeout 600 python -m pytest tests/test_upstream_gpu.py -q -k "fused_layernorm or lanes" > gpurun_out/r2f_pytest.txt 2>&1
tail -3 gpurun_out/r2f_pytest.txt
