set -u
mkdir -p gpurun_out
PYT="python -u -m pytest -q -m gpu --timeout=300 --timeout-method=thread -p no:cacheprovider"
run() { local secs=$1 name=$2; shift 2; timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-400)"; return $rc; }
run 300 g6_test_gemm $PYT tests/test_gpu_gemm.py
run 120 g6_clk_ours python tools/gemm_clocks.py ours
run 120 g6_clk_ours_d8 env PR_GEMM_DEBUG=8 python tools/gemm_clocks.py ours
run 120 g6_clk_ours_d2 env PR_GEMM_DEBUG=2 python tools/gemm_clocks.py ours
run 120 g6_clk_ours_d1 env PR_GEMM_DEBUG=1 python tools/gemm_clocks.py ours
run 120 g6_clk_ours_s5o2 env PR_GEMM_STAGES=5 PR_GEMM_OBUF=2 python tools/gemm_clocks.py ours
run 300 g6_bench_linear python tools/bench_linear.py --out gpurun_out/g6_bench_linear.json
run 300 g6_sasrec $PYT tests/test_gpu_sasrec.py tests/test_gpu_e2e.py tests/test_gpu_ln_loss.py
run 400 g6_bench_tc_graph python bench.py --steps 20 --warmup 5 --no-cpu --graph
run 400 g6_bench_tc_graph_nofuse env PR_FUSE_ACT_BWD=0 python bench.py --steps 20 --warmup 5 --no-cpu --graph
