mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_t_tc|k_bk_tc" -s 2 -c 2 -o gpurun_out/${1:-r02_tiled} -f python tools/quick_bench.py 28 1 2 4194304 tc strict tiled > gpurun_out/ncu_${1:-r02_tiled}.log 2>&1
tail -2 gpurun_out/ncu_${1:-r02_tiled}.log
