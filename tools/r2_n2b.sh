set -u
mkdir -p gpurun_out
run() { local secs=$1 name=$2; shift 2; timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
for ex in nccl p2p; do
  run 420 n2_bench_$ex python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --exchange $ex
  tail -1 gpurun_out/n2_bench_$ex.log > gpurun_out/n2_bench_$ex.json
done
