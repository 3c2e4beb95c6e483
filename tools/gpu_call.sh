set -x
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -4
timeout 300 python tools/time_fused.py 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_final2.json 2> gpurun_out/bench_r01_final2.err; head -c 330 gpurun_out/bench_r01_final2.json; tail -3 gpurun_out/bench_r01_final2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference2.json 2>/dev/null; head -c 300 gpurun_out/bench_r01_reference2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6500 --csv --log-file gpurun_out/launches_r01_final2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu_final2.log 2>&1; tail -c 300 gpurun_out/bench_under_ncu_final2.log
