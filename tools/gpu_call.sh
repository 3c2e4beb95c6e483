timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu --tb=short -k "bf16_stored" 2>&1 | grep -E "^E|assert|passed|failed|Error" | head -8
timeout 900 python bench.py --no-cpu > gpurun_out/bench_r01_final4.json 2> gpurun_out/bench_r01_final4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_final4.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['e2e_bf16_features'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_r01_final4.err
