mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/s8_bench_n2.json 2> gpurun_out/s8_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-300 gpurun_out/s8_bench_n2.json; tail -3 gpurun_out/s8_bench_n2.err | cut -c1-300
