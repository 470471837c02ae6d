export ALG_BENCH_NCELL=${NCELL:-28} ALG_BENCH_NMOL=${NMOL:-40000} ALG_BENCH_NATOMS=${NATOMS:-100000}
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu --config ${CFG:-c2} ${EXTRA} 2>&1 | tail -${TAIL:-6}
