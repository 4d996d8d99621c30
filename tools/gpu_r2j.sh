#!/bin/bash
# round-2 pass j (8 GPUs): the sharded bench with the push gather and with NCCL, then BASELINE config 5 under DDP
mkdir -p gpurun_out
export PYTHONPATH=$PWD/oracle/_ref:$PYTHONPATH
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2j_n8_push.json 2> gpurun_out/r2j_n8_push.err
echo "push rc=$?"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 --gather nccl > gpurun_out/r2j_n8_nccl.json 2> gpurun_out/r2j_n8_nccl.err
echo "nccl rc=$?"
rm -rf /tmp/exp_c5_n8
timeout 600 $TR -m s3prl_b200.run_downstream --synthetic_data --stage_timing -m train -u hubert_base -d ctc \
    -c downstream/ctc/librispeech.yaml -p /tmp/exp_c5_n8 \
    -o "config.runner.total_steps=24,,config.runner.eval_step=100000,,config.runner.save_step=100000,,config.runner.log_step=8" \
    > gpurun_out/r2j_config5_n8.log 2>&1
echo "config5 rc=$?" >> gpurun_out/r2j_config5_n8.log
grep s3b_stage_timing gpurun_out/r2j_config5_n8.log | head -3
