#!/bin/bash
# round-2 first GPU pass: parity tests, lanes A/B at the full batch and at the 8-GPU shard size
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2a_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.txt
for lanes in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --lanes $lanes --no-cpu-baseline > gpurun_out/r2a_c2_l$lanes.json 2> gpurun_out/r2a_c2_l$lanes.err
  timeout 300 python bench.py --steps 20 --warmup 3 --emulate-world 8 --lanes $lanes --no-cpu-baseline > gpurun_out/r2a_shard8_l$lanes.json 2> gpurun_out/r2a_shard8_l$lanes.err
done
tail -5 gpurun_out/r2a_pytest.txt
