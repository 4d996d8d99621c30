#!/bin/bash
# round-2 pass o: f16q8 as the default operand scheme — full parity suite, smoke, bench (default and bf16x3)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2o_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.txt
tail -4 gpurun_out/r2o_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2o_smoke.txt 2>&1; tail -1 gpurun_out/r2o_smoke.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench_c2.json 2> gpurun_out/r2o_bench_c2.err
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_c4.json 2> gpurun_out/r2o_bench_c4.err
S3B_GEMM_SCHEME=bf16x3 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_bench_c2_bf16x3.json 2> gpurun_out/r2o_bench_c2_bf16x3.err
