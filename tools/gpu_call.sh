set -x
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu --tb=short 2>&1 | grep -v Warning | tail -25
DVGR_GAT_FAST=0 timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu --tb=line 2>&1 | tail -4
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "pair or colsum or prep" --tb=short 2>&1 | tail -5
timeout 300 python tools/time_fused.py 2>&1 | tail -2
