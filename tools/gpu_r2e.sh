#!/bin/bash
# round-2 pass e: the f16q8 operand scheme (fp16 product + two e4m3 corrections): block tests, full parity suite, A/B bench
mkdir -p gpurun_out
S3B_GEMM_SCHEME=f16q8 timeout 600 python -m pytest tests/test_blocks_gpu.py -q -s -k "linear" > gpurun_out/r2e_blocks_q8.txt 2>&1
echo "rc=$?" >> gpurun_out/r2e_blocks_q8.txt
grep -c PASSED gpurun_out/r2e_blocks_q8.txt; grep "rel=" gpurun_out/r2e_blocks_q8.txt | head -12; tail -3 gpurun_out/r2e_blocks_q8.txt
S3B_GEMM_SCHEME=f16q8 timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=20 > gpurun_out/r2e_pytest_q8.txt 2>&1
echo "rc=$?" >> gpurun_out/r2e_pytest_q8.txt
tail -5 gpurun_out/r2e_pytest_q8.txt
for sch in f16q8 bf16x3; do
  S3B_GEMM_SCHEME=$sch timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_c2_$sch.json 2> gpurun_out/r2e_c2_$sch.err
done
S3B_GEMM_SCHEME=f16q8 timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_c3_f16q8.json 2> gpurun_out/r2e_c3_f16q8.err
S3B_GEMM_SCHEME=f16q8 timeout 300 python bench.py --steps 20 --warmup 3 --emulate-world 8 --no-cpu-baseline > gpurun_out/r2e_shard8_f16q8.json 2> gpurun_out/r2e_shard8_f16q8.err
