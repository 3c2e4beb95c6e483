set -x
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_s2d.json 2> gpurun_out/bench_r01_s2d.err; tail -c 1800 gpurun_out/bench_r01_s2d.json; tail -5 gpurun_out/bench_r01_s2d.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r01_s2d.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu_s2d.log 2>&1; tail -3 gpurun_out/bench_under_ncu_s2d.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_seq -c 8 -o gpurun_out/prof_r01_lstm_seq python tools/prof_lstm_once.py 2>&1 | tail -5
ls -la gpurun_out | tail -5
