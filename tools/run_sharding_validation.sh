#!/bin/bash
# 2-GPU box: NVLink / NCCL parity tests incl. the stage sharded over both GPUs, then what the sharding costs per step.
mkdir -p gpurun_out
echo "== multigpu tests"; date +%T
timeout 600 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2_pytest_multigpu_sharded.log
echo "== sharded stage K=2"; date +%T
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
  tools/sharded_stage_timing.py > gpurun_out/r2_sharded_stage_k2.json 2> gpurun_out/r2_sharded_stage_k2.err
tail -3 gpurun_out/r2_sharded_stage_k2.err; cat gpurun_out/r2_sharded_stage_k2.json
echo "== unsharded K=1"; date +%T
timeout 300 python tools/sharded_stage_timing.py > gpurun_out/r2_sharded_stage_k1.json 2> gpurun_out/r2_sharded_stage_k1.err
tail -3 gpurun_out/r2_sharded_stage_k1.err; cat gpurun_out/r2_sharded_stage_k1.json
date +%T
