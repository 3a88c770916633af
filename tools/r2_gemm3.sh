set -u
mkdir -p gpurun_out
for cg in 2 1; do for dbg in 0 1 4 2 6; do
  timeout -s KILL 120 env PR_GEMM_CG=$cg PR_GEMM_DEBUG=$dbg python tools/bench_gemm_debug.py 2>&1 | tail -1 | tee -a gpurun_out/g3_debug.jsonl
done; done
timeout -s KILL 120 env PR_GEMM_STAGES=3 python tools/bench_gemm_debug.py 2>&1 | tail -1 | tee -a gpurun_out/g3_debug.jsonl
PYT="python -u -m pytest -q -m gpu --timeout=300 --timeout-method=thread -p no:cacheprovider"
timeout -s KILL 300 $PYT tests/test_gpu_score.py -k "exact" > gpurun_out/g3_score_exact.log 2>&1; echo "score_exact rc=$? $(tail -1 gpurun_out/g3_score_exact.log)"
