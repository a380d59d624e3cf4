cd /root/repo
echo "== tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "cross_entropy or attention" 2>&1 | tail -3
echo "== bench N=1"; date +%T; timeout 1000 python bench.py 2> gpurun_out/r2_bench_n1.err | tail -1 > gpurun_out/r2_bench_xl_n1.json; date +%T; grep -v Warn gpurun_out/r2_bench_n1.err | tail -5; cut -c1-400 gpurun_out/r2_bench_xl_n1.json
echo "== ncu attention bwd"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 6 -c 2 -o gpurun_out/r2_attn_bwd_v4 python tools/ncu_attn.py 1 > gpurun_out/ncu_attn_bwd_v4.log 2>&1; tail -2 gpurun_out/ncu_attn_bwd_v4.log
echo "== XL stage parity"; timeout 600 python -m pytest tests/test_stage_gpu.py -x -q -s -k "benchmark_dims or loss_scale_stress" > gpurun_out/r2_stage_xl2.log 2>&1; grep -E "viol|bwd_fp16=|passed|failed|Error" gpurun_out/r2_stage_xl2.log | cut -c1-220
