set -x
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r01_s2i.json 2> gpurun_out/bench_r01_s2i.err; head -c 330 gpurun_out/bench_r01_s2i.json; tail -5 gpurun_out/bench_r01_s2i.err
timeout 600 python tools/profile_step.py > gpurun_out/profile_step_r01_s2i.txt 2>&1; head -36 gpurun_out/profile_step_r01_s2i.txt | cut -c1-120
