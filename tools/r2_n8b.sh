set -u
mkdir -p gpurun_out
run() { local secs=$1 name=$2; shift 2; timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-200)"; return $rc; }
run 150 q_bench_n8 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu
grep -h '^{"metric"' gpurun_out/q_bench_n8.log | tail -1 > gpurun_out/q_bench_n8.json
run 150 q_bench_n4 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu
grep -h '^{"metric"' gpurun_out/q_bench_n4.log | tail -1 > gpurun_out/q_bench_n4.json
run 150 q_bench_n1 python bench.py --steps 20 --warmup 5 --no-cpu
grep -h '^{"metric"' gpurun_out/q_bench_n1.log | tail -1 > gpurun_out/q_bench_n1.json
