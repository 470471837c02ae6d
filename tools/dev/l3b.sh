timeout 300 python tools/quick_bench.py 30 3 3 2097152 tc strict tiled 2>&1 | grep -E "iter 3|kernel ms" | cut -c1-300
timeout 300 python tools/quick_bench.py 30 3 3 2097152 tc strict fused 2>&1 | grep -E "iter 3" | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_configs.py tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -2
