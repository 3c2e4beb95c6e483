timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu --tb=short -k "early_gradient" 2>&1 | grep -E "^E|assert|passed|failed|Error" | head -12
