#!/bin/bash
# round-2 pass m: final-build evidence — full parity suite, smoke, bench lines of every config, ncu launch list + DRAM
# traffic of one C2 step, ncu --set full of one launch per GEMM site / attention (both operand schemes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/r2m_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_pytest.txt
tail -3 gpurun_out/r2m_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2m_smoke.txt 2>&1; tail -1 gpurun_out/r2m_smoke.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench_c2.json 2> gpurun_out/r2m_bench_c2.err
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_c3.json 2> gpurun_out/r2m_bench_c3.err
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_c4.json 2> gpurun_out/r2m_bench_c4.err
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2m_launches_c2.csv python tools/profile_step.py --steps 1 --warmup 1 --lanes 1 > gpurun_out/r2m_launches_c2.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel|attention_kernel|conv0_apply|posconv_combine" -s 77 -c 8 -f \
    -o gpurun_out/prof_r2m python tools/profile_step.py --steps 1 --warmup 1 --lanes 1 > gpurun_out/r2m_prof.log 2>&1
S3B_GEMM_SCHEME=f16q8 timeout 500 ncu --set full --clock-control none -k regex:"gemm2_kernel" -s 62 -c 6 -f \
    -o gpurun_out/prof_r2m_q8 python tools/profile_step.py --steps 1 --warmup 1 --lanes 1 > gpurun_out/r2m_prof_q8.log 2>&1
