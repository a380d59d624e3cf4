cd /root/repo
export OOB_P2P_TIMEOUT_S=60
echo "== reconfig 2x2 xl"; timeout 500 python bench.py --reconfig --gpus 4 --replicas 2 --model gpt2-xl 2>&1 | grep -v "Warning\|warn\|return func\|^NCCL" | tail -4 | tee gpurun_out/r2_reconfig_xl_2x2_v2.log
echo "== reconfig lone 4 xl"; timeout 500 python bench.py --reconfig --gpus 4 --replicas 1 --model gpt2-xl 2>&1 | grep -v "Warning\|warn\|return func\|^NCCL" | tail -4 | tee gpurun_out/r2_reconfig_xl_lone4_v2.log
echo "== reconfig 2x2 gpt2"; timeout 300 python bench.py --reconfig --gpus 4 --replicas 2 --model gpt2 2>&1 | grep -v "Warning\|warn\|return func\|^NCCL" | tail -4 | tee gpurun_out/r2_reconfig_gpt2_2x2_v2.log
