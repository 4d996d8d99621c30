#!/bin/bash
# round-2 pass k: persistent attention kernel — block tests, full parity suite, same-box A/B; ncu of the fbank / mel kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_blocks_gpu.py -q -s -k "attention" > gpurun_out/r2k_attn_blocks.txt 2>&1
echo "rc=$?" >> gpurun_out/r2k_attn_blocks.txt
grep "rel=\|passed\|failed\|rc=" gpurun_out/r2k_attn_blocks.txt | tail -12
if grep -q "rc=0" gpurun_out/r2k_attn_blocks.txt; then
  timeout 900 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/r2k_pytest.txt 2>&1
  echo "pytest rc=$?" >> gpurun_out/r2k_pytest.txt
  tail -4 gpurun_out/r2k_pytest.txt
  for persist in 1 0; do
    S3B_ATTN_PERSIST=$persist timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_c2_persist$persist.json 2> gpurun_out/r2k_c2_persist$persist.err
    S3B_ATTN_PERSIST=$persist timeout 200 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_c3_persist$persist.json 2> gpurun_out/r2k_c3_persist$persist.err
  done
fi
timeout 300 ncu --set full --clock-control none -k regex:"fbank|stft_mel|mel_cmvn|trimmed|delta" \
    -c 12 -f -o gpurun_out/prof_fbank python tools/profile_small.py > gpurun_out/r2k_prof_fbank.log 2>&1
