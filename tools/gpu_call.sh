timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short 2>&1 | grep -E "^E|assert|passed|failed" | head -12
