cd /root/repo
echo "== attention timing"; timeout 200 python tools/ncu_attn.py 1 2>&1 | tail -2 | tee gpurun_out/r2_attn_time5.log
timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -2
echo "== bench N=1"; timeout 900 python bench.py 2> gpurun_out/r2_bench_n1.err | tail -1 > gpurun_out/r2_bench_xl_n1.json; tail -3 gpurun_out/r2_bench_n1.err; cut -c1-600 gpurun_out/r2_bench_xl_n1.json
echo "== ncu attention"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 9 -c 3 -o gpurun_out/r2_attn_v3 python tools/ncu_attn.py 1 > gpurun_out/ncu_attn_v3.log 2>&1; tail -2 gpurun_out/ncu_attn_v3.log
echo "== ncu gemm fc"; timeout 300 ncu --set full --clock-control none -k regex:gemm_bf16x3 -s 3 -c 1 -o gpurun_out/r2_gemm_fc_pair python tools/ncu_gemm_fc_pair.py > gpurun_out/ncu_gemm_fc.log 2>&1; tail -2 gpurun_out/ncu_gemm_fc.log
