timeout 600 python -m pytest tests/test_gpu_self_bwd.py -q -m gpu 2>&1 | tail -3
