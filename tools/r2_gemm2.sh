set -u
mkdir -p gpurun_out
PYT="python -u -m pytest -q -m gpu --timeout=300 --timeout-method=thread -p no:cacheprovider"
run() { local secs=$1 name=$2; shift 2; timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
run 400 g2_test_gemm $PYT tests/test_gpu_gemm.py
run 300 g2_bench_default python tools/bench_linear.py --out gpurun_out/g2_bench_linear_default.json
run 300 g2_bench_s5o2 env PR_GEMM_STAGES=5 PR_GEMM_OBUF=2 python tools/bench_linear.py --out gpurun_out/g2_bench_linear_s5o2.json
run 300 g2_bench_s4 env PR_GEMM_STAGES=4 python tools/bench_linear.py --out gpurun_out/g2_bench_linear_s4.json
run 900 g2_pytest_all python -m pytest tests -x -q -m gpu
run 400 g2_bench_tc python bench.py --steps 20 --warmup 5 --no-cpu
run 400 g2_bench_cublas env PR_LINEAR=cublas python bench.py --steps 20 --warmup 5 --no-cpu
run 400 g2_bench_tc_graph python bench.py --steps 20 --warmup 5 --no-cpu --graph
run 400 g2_bench_b64_graph python bench.py --batch 64 --steps 200 --warmup 20 --no-cpu --graph
NCU="ncu --set full --clock-control none --import-source on"
run 500 g2_ncu_gemm $NCU -k regex:gemm_tf32 -s 8 -c 2 -o gpurun_out/g2_gemm python tools/bench_linear.py --iters 1
for f in gpurun_out/g2_*.ncu-rep; do [ -f "$f" ] && ncu -i "$f" --page raw --csv > "${f%.ncu-rep}.raw.csv" 2>/dev/null; done
