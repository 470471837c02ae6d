# ncu full capture of the fused kernel (quick_bench box) -> gpurun_out/
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fused_tc -s 2 -c 1 -o gpurun_out/${1:-r02_fused} -f python tools/quick_bench.py 28 ${2:-1} ${3:-2} 2097152 tc strict fused > gpurun_out/ncu_${1:-r02_fused}.log 2>&1
tail -3 gpurun_out/ncu_${1:-r02_fused}.log
