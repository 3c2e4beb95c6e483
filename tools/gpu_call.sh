set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "lstm" --tb=short 2>&1 | grep -v Warning | tail -5
timeout 300 python tools/time_lstm_seq.py 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r01_s2j.json 2> gpurun_out/bench_r01_s2j.err; head -c 330 gpurun_out/bench_r01_s2j.json; tail -5 gpurun_out/bench_r01_s2j.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_seq -c 2 -o gpurun_out/prof_r01_lstm_seq_v2 python tools/prof_lstm_once.py 2>&1 | grep -E "error|timeouts" | tail -3
for k in aux_grad aux_gram; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/prof_r01_$k python tools/time_fused.py 2>&1 | grep -E "error" | tail -2
done
