timeout 300 python tools/quick_bench.py 30 3 3 2097152 tc strict tiled 2>&1 | grep -E "iter 3|kernel ms" | cut -c1-260
