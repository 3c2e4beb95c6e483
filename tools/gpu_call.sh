timeout 900 python -m pytest tests -x -q -m gpu --tb=short 2>&1 | grep -E "^E|assert|passed|failed|Error" | head -8
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e 2>/dev/null | head -c 230; echo
