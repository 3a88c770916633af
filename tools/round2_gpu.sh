#!/bin/bash
# Validation of the code staged without GPU access (score_topk v2 / multicast, long-sequence attention, peer kernels, p2p
# exchange).  Every stage runs under its own timeout so a hung kernel cannot hold the box; logs -> gpurun_out/.
#
#   gpurun --timeout 1500 -- 'bash tools/round2_gpu.sh stage1'            # 1 GPU: staged kernels, parity + A/B timing
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/round2_gpu.sh stage2'   # 2 GPUs: peer-memory exchange vs NCCL
#   gpurun --timeout 1500 -- 'bash tools/round2_gpu.sh configs'           # 1 GPU: BASELINE configs 3, 4, 5
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/round2_gpu.sh stage8'  # 8 GPUs: N=8 bench, both exchanges; config 3 sharded 8-way
#   gpurun --timeout 1800 -- 'bash tools/round2_gpu.sh ncu'               # 1 GPU: ncu --set full of the staged kernels
#
# Order matters: parity first (cheap, tells which variant may become a default), timing after.
set -u
mkdir -p gpurun_out
export PR_EXPERIMENTAL=1
stage=${1:-stage1}
run() {  # run <seconds> <log name> <command...>
  local secs=$1 name=$2; shift 2
  timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-200)"
  grep -E "FAILED|Timeout|rror" "gpurun_out/$name.log" | head -6
  return $rc
}
PYT="python -u -m pytest -q -m gpu --timeout=120 --timeout-method=thread -p no:cacheprovider"

if [ "$stage" = stage1 ]; then
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
  # 1. SIMT kernels whose logic is already pinned by the CPU emulation (lowest risk first)
  run 300 r2_attn_long   $PYT tests/test_gpu_attn_long.py
  run 300 r2_vit         $PYT tests/test_gpu_vit.py
  run 300 r2_peer        $PYT tests/test_gpu_peer.py
  # 2. score_topk variants: v2 epilogue (tune 16), then + cluster multicast (tune 48); each variant in its own process
  run 300 r2_score_v2    $PYT tests/test_gpu_score.py -k "v2 and not mcast"
  run 300 r2_score_mcast $PYT tests/test_gpu_score.py -k "v2_mcast"
  run 300 r2_linear      $PYT tests/test_gpu_linear.py
  run 300 r2_score_ce    $PYT tests/test_gpu_score.py -k "score_ce and v1"
  run 300 r2_score_f16   $PYT tests/test_gpu_score.py -k "f16 and v1"
  run 300 r2_score_f16mc $PYT tests/test_gpu_score.py -k "f16 and v2_mcast"
  # 3. timing (only meaningful if the parity runs above passed)
  run 300 r2_bench_attn_long python tools/bench_attn_long.py
  run 300 r2_bench_score env SCORE_TUNES=0,16,48 SCORE_F16=1 python tools/bench_score.py
  run 300 r2_bench_score_f16mc env PR_TUNE=$((9 | 32)) SCORE_F16=1 python tools/bench_score.py
  run 300 r2_bench_score_mc_cl4 env PR_SCORE_CLUSTER=4 SCORE_TUNES=48 SCORE_F16=1 PR_TUNE=$((9 | 32)) python tools/bench_score.py
  run 300 r2_bench_score_f16mc_ares env PR_TUNE=$((9 | 32 | 128)) SCORE_F16=1 python tools/bench_score.py
  cp gpurun_out/bench_score.json gpurun_out/r2_bench_score.json 2>/dev/null
  # 4. the whole default suite + bench, as the driver runs them
  run 900 r2_pytest_default env -u PR_EXPERIMENTAL python -m pytest tests -x -q -m gpu
  run 600 r2_bench_n1 python bench.py --steps 20 --warmup 5
  run 600 r2_plan_ahead_parity env PR_PLAN_AHEAD=1 python -m pytest tests/test_gpu_sasrec.py tests/test_gpu_e2e.py -x -q -m gpu
  run 600 r2_bench_n1_plan_ahead env PR_PLAN_AHEAD=1 python bench.py --steps 20 --warmup 5 --no-cpu   # scatter plan overlapped with the forward
  # CUDA-graph replay of the step: dropout-free parity on the default build, then a -DPR_SEED_DEV build for the shipped dropout
  run 300 r2_graph_p0 $PYT tests/test_gpu_graph.py
  run 300 r2_bench_b64_eager python bench.py --batch 64 --steps 200 --warmup 20 --no-cpu
  cp pixelrec_b200/libpixelrec_b200.so /tmp/libpixelrec_b200.default.so
  run 600 r2_build_seeddev env PR_BUILD_DEFS=-DPR_SEED_DEV python -m pixelrec_b200.build --force
  run 300 r2_graph_p01 $PYT tests/test_gpu_graph.py
  run 300 r2_bench_b64_graph python bench.py --batch 64 --steps 200 --warmup 20 --no-cpu --graph
  run 300 r2_bench_b4096_graph python bench.py --steps 20 --warmup 5 --no-cpu --graph
  cp /tmp/libpixelrec_b200.default.so pixelrec_b200/libpixelrec_b200.so      # back to the default binary for the stages below
  tail -1 gpurun_out/r2_bench_n1.log > gpurun_out/r2_bench_n1.json
elif [ "$stage" = ncu ]; then
  # one full capture per staged kernel that passed stage1 (never a bench number: ncu replays every kernel ~40 times)
  NCU="ncu --set full --clock-control none --import-source on"
  run 600 r2_ncu_score_v2 env SCORE_TUNES=16 $NCU -k regex:score_topk2 -c 1 -o gpurun_out/r2_score_v2 python tools/bench_score.py
  run 600 r2_ncu_score_mc env SCORE_TUNES=48 $NCU -k regex:score_topk2 -s 6 -c 1 -o gpurun_out/r2_score_mcast python tools/bench_score.py
  run 600 r2_ncu_score_f16 env PR_TUNE=$((9 | 32 | 128)) SCORE_F16=1 $NCU -k regex:score_topk2 -c 1 -o gpurun_out/r2_score_f16 python tools/bench_score.py
  run 600 r2_ncu_attn_long $NCU -k regex:attn_long -c 3 -o gpurun_out/r2_attn_long python tools/bench_attn_long.py --images 64
  for f in gpurun_out/r2_*.ncu-rep; do [ -f "$f" ] && ncu -i "$f" --page raw --csv > "${f%.ncu-rep}.raw.csv" 2>/dev/null; done
elif [ "$stage" = configs ]; then
  # the BASELINE.json configurations bench.py does not cover, one GPU (c3 also fits one GPU: 3.35 GB table + Adam state)
  for c in c5 c3; do
    run 420 r2_cfg_$c python tools/bench_configs.py --config $c
    tail -1 gpurun_out/r2_cfg_$c.log > gpurun_out/r2_cfg_$c.json
  done
  run 600 r2_cfg_c4_p32 python tools/bench_configs.py --config c4 --patch 32 --steps 5
  run 600 r2_cfg_c4_p16 python tools/bench_configs.py --config c4 --patch 16 --steps 5
elif [ "$stage" = stage2 ]; then
  run 600 r2_dist $PYT tests/test_gpu_dist.py
  for ex in nccl p2p; do
    run 420 r2_bench_n2_$ex python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --exchange $ex
    tail -1 gpurun_out/r2_bench_n2_$ex.log > gpurun_out/r2_bench_n2_$ex.json
  done
elif [ "$stage" = stage8 ]; then
  # 8 GPUs (charged 8x): only after stage2 has confirmed the peer exchange at N=2
  for ex in p2p nccl; do
    run 420 r2_bench_n8_$ex python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --gpus 8 --steps 20 --warmup 5 --exchange $ex
    tail -1 gpurun_out/r2_bench_n8_$ex.log > gpurun_out/r2_bench_n8_$ex.json
  done
  run 420 r2_cfg_c3_n8 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
      tools/bench_configs.py --config c3 --exchange p2p
fi
exit 0

# ---- stages j / k (second half of round 2): the calls behind profiles/r02j_* and r02k_*, one gpurun call per stage
if [ "$stage" = rowk ]; then            # segment reduce / plan / LayerNorm / loss kernels: parity, A/B, bench, ncu
  run 700 k_pytest_all   python -m pytest tests -m gpu -q --maxfail=40
  run 300 k_bench_n1     python bench.py
  run 300 k_bench_noz    env PR_FUSE_LN_Z=0 python bench.py --no-cpu              # dropout + residual NOT in the GEMM epilogue
  run 300 k_bench_ln1row env PR_TUNE=345 python bench.py --no-cpu                  # LayerNorm forward, one row per warp
  for v in 0 4 8 1; do                                                             # ring v2 (3 / 4 stages), ring v1 (register queue), v1 with 8 KiB stages
    run 120 k_scatter_v$v env PR_SCATTER_VARIANT=$v python tools/bench_scatter.py --json gpurun_out/k_scatter_v$v.json
  done
  run 300 k_ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/k_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu
  run 300 k_ncu_rowk ncu --set full --clock-control none --import-source on \
      -k regex:"scatter_add_rows_ring|add_ln_fwd_rows2|bpr_fwd|bpr_bwd" -c 8 -o gpurun_out/k_rowkernels \
      python bench.py --steps 1 --warmup 1 --no-cpu --no-graph
fi
if [ "$stage" = dist2 ]; then           # gpurun --gpus 2: exchanges, graph replay, sharded evaluation; N=2 bench
  run 500 k_pytest_dist python -m pytest tests/test_gpu_dist.py -m gpu -q
  run 240 k_bench_n2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu
fi
