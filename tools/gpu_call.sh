set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gat or pair" --tb=short 2>&1 | tail -12
timeout 300 python tools/time_fused.py 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_attn_fwd_mma -s 3 -c 1 -o gpurun_out/prof_r01_gat_attn_fwd_mma_v2 python tools/time_fused.py 2>&1 | grep -E "error" | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_attn_bwd_mma -s 3 -c 1 -o gpurun_out/prof_r01_gat_attn_bwd_mma_v2 python tools/time_fused.py 2>&1 | grep -E "error" | tail -2
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r01_s2f.json 2> gpurun_out/bench_r01_s2f.err; tail -c 900 gpurun_out/bench_r01_s2f.json; tail -5 gpurun_out/bench_r01_s2f.err
