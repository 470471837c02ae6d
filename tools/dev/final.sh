mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final_c2.json 2> gpurun_out/bench_final_c2.err
tail -c 600 gpurun_out/bench_final_c2.json; tail -3 gpurun_out/bench_final_c2.err
