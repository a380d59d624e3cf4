cd /root/repo
export OOB_P2P_TIMEOUT_S=60
echo "== multigpu tests"; timeout 400 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2_multigpu_test.log
echo "== config 4 shape (2x2) gpt2"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 4 --replicas 2 --model gpt2 --steps 3 --warmup 2 --cpu-baseline 0 2>&1 | tail -4 | tee gpurun_out/r2_bench_gpt2_dp2pp2.log
echo "== reconfig 2x2 gpt2"; timeout 400 python bench.py --reconfig --gpus 4 --replicas 2 --model gpt2 2>&1 | tail -6 | tee gpurun_out/r2_reconfig_gpt2_2x2.log
echo "== reconfig lone 4 gpt2"; timeout 400 python bench.py --reconfig --gpus 4 --replicas 1 --model gpt2 2>&1 | tail -6 | tee gpurun_out/r2_reconfig_gpt2_lone4.log
echo "== reconfig 2x2 xl"; timeout 500 python bench.py --reconfig --gpus 4 --replicas 2 --model gpt2-xl 2>&1 | tail -6 | tee gpurun_out/r2_reconfig_xl_2x2.log
echo "== reconfig lone 4 xl"; timeout 500 python bench.py --reconfig --gpus 4 --replicas 1 --model gpt2-xl 2>&1 | tail -6 | tee gpurun_out/r2_reconfig_xl_lone4.log
