#!/bin/bash
# 4-GPU box: bench.py as the driver launches it at N=4 (small model to keep the box time short), including the embedded
# reconfiguration measurement: 2 x 2 -> 2 + 1 and a lone 4-stage pipeline -> 3 stages (peer shadows, two-phase commit).
mkdir -p gpurun_out
date +%T
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 \
  bench.py --gpus 4 --model gpt2 --steps 3 --warmup 3 \
  > gpurun_out/r2_bench_gpt2_n4_with_reconfig.json 2> gpurun_out/r2_bench_gpt2_n4_with_reconfig.err
echo "rc=$?"; tail -3 gpurun_out/r2_bench_gpt2_n4_with_reconfig.err; cat gpurun_out/r2_bench_gpt2_n4_with_reconfig.json
date +%T
