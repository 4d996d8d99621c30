#!/bin/bash
# round-2 pass n: final build (tile-picker fix, DistilHuBERT) — full parity suite, smoke, bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2n_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.txt
tail -6 gpurun_out/r2n_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2n_smoke.txt 2>&1; tail -1 gpurun_out/r2n_smoke.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_c2.json 2> gpurun_out/r2n_bench_c2.err
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_c3.json 2> gpurun_out/r2n_bench_c3.err
S3B_GEMM_SCHEME=f16q8 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_c2_f16q8.json 2> gpurun_out/r2n_bench_c2_f16q8.err
S3B_GEMM_SCHEME=f16q8 timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_c3_f16q8.json 2> gpurun_out/r2n_bench_c3_f16q8.err
timeout 300 python bench.py --steps 20 --warmup 5 --emulate-world 8 --no-cpu-baseline > gpurun_out/r2n_bench_shard8.json 2> gpurun_out/r2n_bench_shard8.err
