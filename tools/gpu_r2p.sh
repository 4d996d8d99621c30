#!/bin/bash
# round-2 pass p: data2vec audio (conv-block positional encoder) parity, full suite, smoke, final bench, shard-8 lanes A/B
mkdir -p gpurun_out
timeout 240 python tools/check_data2vec.py > gpurun_out/r2p_data2vec.txt 2>&1; tail -6 gpurun_out/r2p_data2vec.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2p_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.txt
tail -4 gpurun_out/r2p_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2p_smoke.txt 2>&1; tail -1 gpurun_out/r2p_smoke.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench_c2.json 2> gpurun_out/r2p_bench_c2.err
timeout 200 python bench.py --emulate-world 8 --lanes 1 --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/r2p_bench_shard8_lanes1.json 2> gpurun_out/r2p_bench_shard8_lanes1.err
timeout 200 python bench.py --emulate-world 8 --lanes 2 --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/r2p_bench_shard8_lanes2.json 2> gpurun_out/r2p_bench_shard8_lanes2.err
