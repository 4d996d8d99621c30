#!/bin/bash
# round-2 pass d: parity tests with the fused LayerNorm + new tile picker; A/B of fused LN; shard-size step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2d_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.txt
tail -4 gpurun_out/r2d_pytest.txt
for fuse in 1 0; do
  S3B_FUSE_LN=$fuse timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_c2_fuse$fuse.json 2> gpurun_out/r2d_c2_fuse$fuse.err
  S3B_FUSE_LN=$fuse timeout 300 python bench.py --steps 20 --warmup 3 --emulate-world 8 --no-cpu-baseline > gpurun_out/r2d_shard8_fuse$fuse.json 2> gpurun_out/r2d_shard8_fuse$fuse.err
done
timeout 300 python bench.py --steps 20 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/r2d_c2_l1.json 2> gpurun_out/r2d_c2_l1.err
timeout 300 python bench.py --steps 20 --warmup 3 --emulate-world 8 --lanes 1 --no-cpu-baseline > gpurun_out/r2d_shard8_l1.json 2> gpurun_out/r2d_shard8_l1.err
timeout 300 python bench.py --steps 20 --warmup 3 --emulate-world 4 --no-cpu-baseline > gpurun_out/r2d_shard4.json 2> gpurun_out/r2d_shard4.err
timeout 300 python bench.py --steps 20 --warmup 3 --emulate-world 2 --no-cpu-baseline > gpurun_out/r2d_shard2.json 2> gpurun_out/r2d_shard2.err
