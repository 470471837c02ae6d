mkdir -p gpurun_out
timeout 320 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --config c5 --steps 3 --warmup 3 > gpurun_out/bench_c5_n4.json 2> gpurun_out/bench_c5_n4.err
tail -c 300 gpurun_out/bench_c5_n4.json; grep -E "rank 0|Error|error" gpurun_out/bench_c5_n4.err | tail -4
