timeout 300 python tools/quick_bench.py 30 3 3 2097152 tc strict tiled 2>&1 | grep -E "iter 3|kernel ms"
timeout 300 python tools/quick_bench.py 30 2 2 2097152 tc strict tiled 2>&1 | grep -E "iter 3|kernel ms"
timeout 300 python tools/quick_bench.py 30 2 2 2097152 generic 2>&1 | grep -E "iter 3"
