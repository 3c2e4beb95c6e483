set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "lstm" --tb=short 2>&1 | tail -25
timeout 300 python tools/time_lstm_seq.py 2>&1 | tail -6
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -60
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r01_s2b.json 2> gpurun_out/bench_r01_s2b.err; tail -c 1500 gpurun_out/bench_r01_s2b.json; tail -5 gpurun_out/bench_r01_s2b.err
timeout 600 python tools/profile_step.py > gpurun_out/profile_step_r01_s2b.txt 2>&1; head -30 gpurun_out/profile_step_r01_s2b.txt
