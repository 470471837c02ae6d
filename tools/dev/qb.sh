timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -4
for b in ${BATCHES:-8}; do echo "== batch $b"; ALG_FUSED_BATCH=$b timeout 200 python tools/quick_bench.py ${NCELL:-40} ${LMAX:-1} ${NLAYER:-2} 2097152 tc strict fused 2>&1 | tail -3 | head -2; done
timeout 200 python tools/quick_bench.py ${NCELL:-40} ${LMAX:-1} ${NLAYER:-2} 2097152 tc strict tiled 2>&1 | tail -3 | head -2
