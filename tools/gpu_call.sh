set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "lstm" --tb=short 2>&1 | tail -15
timeout 300 python tools/time_lstm_seq.py 2>&1 | tail -6
DVGR_LSTM_PREFETCH=0 timeout 300 python tools/time_lstm_seq.py 2>&1 | tail -6
timeout 600 python tools/diag_grad_norms.py g2_B3_N20_U3 24 8 2 2>&1 | tail -30
DVGR_LSTM_SEQ=0 timeout 600 python tools/diag_grad_norms.py g2_B3_N20_U3 2>&1 | tail -18
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r01_s2c.json 2> gpurun_out/bench_r01_s2c.err; tail -c 1500 gpurun_out/bench_r01_s2c.json; tail -5 gpurun_out/bench_r01_s2c.err
