#!/bin/bash
# one-off A/B run on the GPU box: LN variants parity + timing, wgrad GEMM forms, ncu captures of the row kernels
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_ln_loss.py -q -m gpu --timeout=60 --timeout-method=thread > gpurun_out/h_ln_tests.log 2>&1
echo "ln tests rc=$?"; tail -4 gpurun_out/h_ln_tests.log
timeout 120 python tools/bench_rowkernels.py --gemm --json gpurun_out/h_rowk.json > gpurun_out/h_rowk.log 2>&1
echo "rowk rc=$?"; grep -E "^tune|^act" gpurun_out/h_rowk.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"add_ln_bwd|act_bwd_bias|act_fwd" -c 8 -o gpurun_out/h_rowk_ncu \
   python tools/bench_rowkernels.py --tunes 0,1 --once > gpurun_out/h_ncu1.log 2>&1
echo "ncu1 rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"attn_tc" -c 4 -o gpurun_out/h_attn_ncu \
   python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/h_ncu2.log 2>&1
echo "ncu2 rc=$?"
for r in h_rowk_ncu h_attn_ncu; do
  [ -f gpurun_out/$r.ncu-rep ] && ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
