#!/bin/bash
# round-2 pass g (2 GPUs): fused weighted-sum + peer-memory push all-gather vs NCCL: bit-exact check, then the bench at N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tools/check_push_gather.py > gpurun_out/r2g_check_push.txt 2>&1
echo "rc=$?" >> gpurun_out/r2g_check_push.txt
tail -6 gpurun_out/r2g_check_push.txt
for mode in push nccl; do
  timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --gather $mode > gpurun_out/r2g_n2_$mode.json 2> gpurun_out/r2g_n2_$mode.err
  echo "bench $mode rc=$?"
done
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_n2_auto.json 2> gpurun_out/r2g_n2_auto.err
