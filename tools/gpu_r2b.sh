#!/bin/bash
# round-2 second GPU pass: launch list at the 8-GPU shard size, ncu --set full of the small kernel families,
# bench lines of the other BASELINE configs, config 5 (Runner.train through the launcher) on one GPU
mkdir -p gpurun_out
export PYTHONPATH=$PWD/oracle/_ref:$PYTHONPATH
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_shard8.csv \
    python tools/profile_step.py --batch 4 --steps 1 --warmup 1 --lanes 1 > gpurun_out/r2b_launches_shard8.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"layernorm_kernel|weighted_sum_kernel|wavlm_gate|fbank|stft_mel|mel_cmvn|conv0_moments|wav_pack|posconv_combine|trimmed" \
    -s 30 -c 40 -f -o gpurun_out/prof_small python tools/profile_small.py > gpurun_out/r2b_prof_small.log 2>&1
for cfg in c3 c3_ll60k c4 c1_fbank; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/r2b_bench_$cfg.json 2> gpurun_out/r2b_bench_$cfg.err
done
timeout 600 python bench.py --steps 20 --warmup 3 --emulate-world 8 --lanes 2 --no-cpu-baseline > gpurun_out/r2b_shard8_l2.json 2> gpurun_out/r2b_shard8_l2.err
rm -rf /tmp/exp_c5
timeout 900 python -m s3prl_b200.run_downstream --synthetic_data --stage_timing -m train -u hubert_base -d ctc \
    -c downstream/ctc/librispeech.yaml -p /tmp/exp_c5 \
    -o "config.runner.total_steps=24,,config.runner.eval_step=100000,,config.runner.save_step=100000,,config.runner.log_step=8" \
    > gpurun_out/r2b_config5_n1.log 2>&1
echo "config5 rc=$?" >> gpurun_out/r2b_config5_n1.log
grep s3b_stage_timing gpurun_out/r2b_config5_n1.log
