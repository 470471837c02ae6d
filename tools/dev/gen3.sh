timeout 600 python -m pytest tests/test_gpu_generic.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python tools/quick_bench.py 40 1 2 2097152 generic 2>&1 | grep -E "iter 3"
ALG_WIDTHS=128,64,128,2,32 timeout 300 python tools/quick_bench.py 30 3 3 2097152 generic 2>&1 | grep -E "iter 3"
ALG_GENERIC_GEMM=ffma ALG_WIDTHS=128,64,128,2,32 timeout 300 python tools/quick_bench.py 30 3 3 2097152 generic 2>&1 | grep -E "iter 3"
timeout 300 python tools/quick_bench.py 30 3 3 2097152 generic 2>&1 | grep -E "iter 3"
