set -u
mkdir -p gpurun_out
run() { local secs=$1 name=$2; shift 2; timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run 300 n8_bench_p2p $TR --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --exchange p2p
tail -1 gpurun_out/n8_bench_p2p.log > gpurun_out/n8_bench_p2p.json
run 300 n8_bench_nccl $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --exchange nccl
tail -1 gpurun_out/n8_bench_nccl.log > gpurun_out/n8_bench_nccl.json
run 300 n8_cfg_c3 $TR --master-port 29514 tools/bench_configs.py --config c3 --exchange p2p
tail -1 gpurun_out/n8_cfg_c3.log > gpurun_out/n8_cfg_c3.json
run 300 n8_cfg_c4 $TR --master-port 29515 tools/bench_configs.py --config c4 --patch 32 --steps 5
tail -1 gpurun_out/n8_cfg_c4.log > gpurun_out/n8_cfg_c4.json
run 300 n4_bench_p2p python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu --exchange p2p
tail -1 gpurun_out/n4_bench_p2p.log > gpurun_out/n4_bench_p2p.json
