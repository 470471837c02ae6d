mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_bk_tc" -s 2 -c 2 -o gpurun_out/r02_l3_bk -f python tools/quick_bench.py 16 3 3 4194304 tc strict tiled > gpurun_out/ncu_l3.log 2>&1
tail -2 gpurun_out/ncu_l3.log | cut -c1-200
ls -la gpurun_out/r02_l3_bk.ncu-rep
