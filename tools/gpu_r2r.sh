#!/bin/bash
# round-2 pass r: last sanity run of the committed build (lane rule + s3b_default_lanes): smoke, a parity subset, shard-8 bench with the default lane choice
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2r_smoke.txt 2>&1; tail -1 gpurun_out/r2r_smoke.txt
timeout 150 python -m pytest tests -m gpu -q -x -k "lanes_are_bit or (golden and (hubert_base or data2vec_base)) or ragged or shard" > gpurun_out/r2r_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2r_pytest.txt; tail -3 gpurun_out/r2r_pytest.txt
timeout 100 python bench.py --emulate-world 8 --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/r2r_bench_shard8_default.json 2> gpurun_out/r2r_bench_shard8_default.err
python -c "
import json;d=json.loads(open('gpurun_out/r2r_bench_shard8_default.json').read().strip().splitlines()[-1]);print('shard8 default lanes',d['config'].get('lanes'),d['ms_per_step'])"
