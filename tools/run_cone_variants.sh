timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/cone_time.py 2>&1 | tail -6
