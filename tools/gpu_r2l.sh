#!/bin/bash
# round-2 pass l: CUDA-graph replay of the forward — bit-identity test, A/B at the 8-GPU shard size and at the full batch
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_upstream_gpu.py -q -k "graph_replay" > gpurun_out/r2l_graph_test.txt 2>&1
echo "rc=$?" >> gpurun_out/r2l_graph_test.txt
tail -15 gpurun_out/r2l_graph_test.txt
for gr in 1 0; do
  S3B_GRAPHS=$gr timeout 200 python bench.py --steps 30 --warmup 8 --emulate-world 8 --no-cpu-baseline > gpurun_out/r2l_shard8_graphs$gr.json 2> gpurun_out/r2l_shard8_graphs$gr.err
  S3B_GRAPHS=$gr timeout 200 python bench.py --steps 20 --warmup 8 --no-cpu-baseline > gpurun_out/r2l_c2_graphs$gr.json 2> gpurun_out/r2l_c2_graphs$gr.err
done
S3B_GRAPHS=1 timeout 200 python bench.py --steps 30 --warmup 8 --emulate-world 4 --no-cpu-baseline > gpurun_out/r2l_shard4_graphs1.json 2> gpurun_out/r2l_shard4_graphs1.err
S3B_GRAPHS=0 timeout 200 python bench.py --steps 30 --warmup 8 --emulate-world 4 --no-cpu-baseline > gpurun_out/r2l_shard4_graphs0.json 2> gpurun_out/r2l_shard4_graphs0.err
