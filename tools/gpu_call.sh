set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "lstm" --tb=short 2>&1 | grep -v Warning | tail -15
timeout 300 python tools/time_lstm_seq.py 2>&1 | tail -5
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r01_s2g.json 2> gpurun_out/bench_r01_s2g.err; tail -c 900 gpurun_out/bench_r01_s2g.json; tail -5 gpurun_out/bench_r01_s2g.err
timeout 600 python tools/profile_step.py > gpurun_out/profile_step_r01_s2g.txt 2>&1; head -24 gpurun_out/profile_step_r01_s2g.txt
