#!/bin/bash
# 2-GPU box: the sharded-stage NCCL test, then bench.py under torchrun with the embedded reconfiguration measurement
# (forced on at N=2 to exercise its mechanics: 2 replicas x 1 stage lose a rank).
mkdir -p gpurun_out
echo "== multigpu tests"; date +%T
timeout 400 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -6 | tee gpurun_out/r2_pytest_multigpu_sharded.log
echo "== bench gpt2 N=2 + embedded reconfiguration"; date +%T
OOB_BENCH_RECONFIG_MIN_GPUS=2 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
  --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --model gpt2 --steps 3 --warmup 3 \
  > gpurun_out/r2_bench_gpt2_n2_with_reconfig.json 2> gpurun_out/r2_bench_gpt2_n2_with_reconfig.err
echo "rc=$?"; tail -3 gpurun_out/r2_bench_gpt2_n2_with_reconfig.err; cat gpurun_out/r2_bench_gpt2_n2_with_reconfig.json
date +%T
nvidia-smi --query-gpu=memory.used --format=csv
