cd /root/repo
echo "== full gpu tests"; date +%T; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_gputest_final.log; date +%T
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== cpu arm N=1"; timeout 700 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2> gpurun_out/r2_cpu_arm_n1.err | tail -1 | cut -c1-1200 | tee gpurun_out/r2_cpu_arm_n1.json; grep "cpu arm" gpurun_out/r2_cpu_arm_n1.err | tail -6
echo "== cpu arm N=8"; timeout 700 python bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2> gpurun_out/r2_cpu_arm_n8.err | tail -1 | cut -c1-1200 | tee gpurun_out/r2_cpu_arm_n8.json; grep "cpu arm" gpurun_out/r2_cpu_arm_n8.err | tail -6
nproc; free -g | head -2; cat /sys/fs/cgroup/cpu.max 2>/dev/null
