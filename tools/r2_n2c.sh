set -u
mkdir -p gpurun_out
PYT="python -u -m pytest -q -m gpu --timeout=600 --timeout-method=thread -p no:cacheprovider"
run() { local secs=$1 name=$2; shift 2; timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
run 600 p2_dist $PYT tests/test_gpu_dist.py tests/test_gpu_peer.py
run 300 p2_ln $PYT tests/test_gpu_ln_loss.py -k colsum_rows
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run 420 p2_bench_graph $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu
tail -1 gpurun_out/p2_bench_graph.log > gpurun_out/p2_bench_graph.json
run 420 p2_bench_eager $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-graph
tail -1 gpurun_out/p2_bench_eager.log > gpurun_out/p2_bench_eager.json
run 420 p2_bench_eager_ncclbar env PR_P2P_BARRIER=nccl $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-graph
run 420 p2_bench_n1 python bench.py --steps 20 --warmup 5 --no-cpu
