timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E|assert|passed|failed" | head -12
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e 2>/dev/null | head -c 230; echo
timeout 600 python tools/profile_step.py > gpurun_out/profile_step_r01_s2k.txt 2>&1; head -22 gpurun_out/profile_step_r01_s2k.txt | tail -21 | cut -c1-110
