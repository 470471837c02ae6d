timeout 300 python -m pytest tests/test_gpu_neigh.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/dev/neigh_time.py 63 2>&1 | tail -6
