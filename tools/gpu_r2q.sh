#!/bin/bash
# round-2 pass q: one lane vs two at the shard sizes of the 2- and 4-GPU runs (16 / 8 utterances of 10 s per rank)
mkdir -p gpurun_out
for w in 2 4; do
  for l in 1 2; do
    timeout 200 python bench.py --emulate-world $w --lanes $l --steps 30 --warmup 8 --no-cpu-baseline \
      > gpurun_out/r2q_bench_shard${w}_lanes${l}.json 2> gpurun_out/r2q_bench_shard${w}_lanes${l}.err
    python -c "
import json;d=json.loads(open('gpurun_out/r2q_bench_shard${w}_lanes${l}.json').read().strip().splitlines()[-1]);print('world',$w,'lanes',$l,d['ms_per_step'],d['clocks']['sm_mhz'])"
  done
done
