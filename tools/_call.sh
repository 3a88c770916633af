mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/s4_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s4_pytest_all.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/s4_bench_n1.json 2> gpurun_out/s4_bench_n1.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/s4_bench_n1.json
PR_TUNE=857 timeout 300 python bench.py --no-cpu > gpurun_out/s4_bench_n1_rows2.json 2> gpurun_out/s4_bench_n1_rows2.err; echo "bench rows2 rc=$?"; cut -c1-260 gpurun_out/s4_bench_n1_rows2.json
for v in 0 1 2 3; do PR_SCATTER_VARIANT=$v timeout 120 python tools/bench_scatter.py --json gpurun_out/s4_scatter_v$v.json > gpurun_out/s4_scatter_v$v.log 2>&1; echo "scatter v$v rc=$?"; grep -h "C2_B4096\|C3_B1024\|gather_distinct" gpurun_out/s4_scatter_v$v.log | cut -c1-420; done
