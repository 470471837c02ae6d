mkdir -p gpurun_out
timeout 900 python bench.py --config c5h > gpurun_out/bench_c5h.json 2> gpurun_out/bench_c5h.err
tail -c 1500 gpurun_out/bench_c5h.json; tail -5 gpurun_out/bench_c5h.err
