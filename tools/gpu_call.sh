for i in 1 2 3 4; do timeout 300 python -m pytest tests/test_engine_gpu.py -x -q -m gpu --tb=short -k graph_replay_matches 2>&1 | grep -E "assert|Error|passed|failed" | head -6; done
