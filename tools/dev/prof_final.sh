mkdir -p gpurun_out
export ALG_BENCH_NCELL=40
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_tc -s 3 -c 1 -o gpurun_out/r02_ncu_fused_c2 -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_fused.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_f0_tc|k_fk_tc|k_t_tc|k_bk_tc|k_b0_tc" -s 10 -c 5 -o gpurun_out/r02_ncu_tiled_c2 -f python bench.py --steps 1 --warmup 3 --no-cpu --pipeline tiled --chunk-edges 8388608 > gpurun_out/ncu_tiled.log 2>&1
tail -2 gpurun_out/ncu_fused.log | cut -c1-150; tail -2 gpurun_out/ncu_tiled.log | cut -c1-150
