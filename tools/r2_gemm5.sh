set -u
mkdir -p gpurun_out
PYT="python -u -m pytest -q -m gpu --timeout=300 --timeout-method=thread -p no:cacheprovider"
run() { local secs=$1 name=$2; shift 2; timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-400)"; return $rc; }
run 300 g5_test_gemm $PYT tests/test_gpu_gemm.py
run 120 g5_clk_cublas python tools/gemm_clocks.py cublas
run 120 g5_clk_ours python tools/gemm_clocks.py ours
run 120 g5_clk_ours_d6 env PR_GEMM_DEBUG=6 python tools/gemm_clocks.py ours
run 300 g5_bench_linear python tools/bench_linear.py --out gpurun_out/g5_bench_linear.json
run 300 g5_score $PYT tests/test_gpu_score.py -k "exact"
run 400 g5_bench_tc python bench.py --steps 20 --warmup 5 --no-cpu
run 400 g5_bench_tc_graph python bench.py --steps 20 --warmup 5 --no-cpu --graph
run 400 g5_bench_cublas_graph env PR_LINEAR=cublas python bench.py --steps 20 --warmup 5 --no-cpu --graph
