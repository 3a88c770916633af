set -u
mkdir -p gpurun_out
PYT="python -u -m pytest -q -m gpu --timeout=90 --timeout-method=thread -p no:cacheprovider"
run() { local secs=$1 name=$2; shift 2; timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
run 200 g1_diag_cg1 env PR_GEMM_CG=1 python tools/diag_gemm.py
run 200 g1_diag_cg2 python tools/diag_gemm.py
run 400 g1_test_cg1 env PR_GEMM_CG=1 $PYT tests/test_gpu_gemm.py
run 400 g1_test_cg2 $PYT tests/test_gpu_gemm.py
run 300 g1_bench_cg2 python tools/bench_linear.py --out gpurun_out/g1_bench_linear_cg2.json
run 300 g1_bench_cg1 env PR_GEMM_CG=1 python tools/bench_linear.py --out gpurun_out/g1_bench_linear_cg1.json
run 300 g1_graph $PYT tests/test_gpu_graph.py
run 300 g1_sasrec $PYT tests/test_gpu_sasrec.py
