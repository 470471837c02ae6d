timeout 500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_halo.py tests/test_gpu_tc.py tests/test_gpu_configs.py tests/test_gpu_pair_style.py -m gpu -x -q 2>&1 | tail -6
bash tools/dev/qb.sh 2>&1 | tail -5
timeout 200 python tools/quick_bench.py 40 1 2 2097152 tc strict auto 2>&1 | tail -3
