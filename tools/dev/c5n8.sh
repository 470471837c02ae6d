mkdir -p gpurun_out
timeout 95 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config c5 --steps 3 --warmup 3 --lean > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err
echo rc=$?
tail -c 300 gpurun_out/bench_c5_n8.json; grep -E "rank 0|Error|error" gpurun_out/bench_c5_n8.err | tail -4
