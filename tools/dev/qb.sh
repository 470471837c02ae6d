python -m pytest tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -8
for b in 1 4 8 16; do echo "== batch $b"; ALG_FUSED_BATCH=$b python tools/quick_bench.py 40 1 2 2097152 tc strict fused 2>&1 | tail -3 | head -2; done
python tools/quick_bench.py 40 1 2 2097152 tc strict tiled 2>&1 | tail -3 | head -2
