mkdir -p gpurun_out
timeout 1500 python bench.py --config c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
tail -c 400 gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err
