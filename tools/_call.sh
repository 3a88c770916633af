mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rows.py tests/test_gpu_e2e.py tests/test_gpu_sasrec.py -m gpu -x -q > gpurun_out/s1_pytest_rows.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s1_pytest_rows.log
timeout 200 python tools/bench_scatter.py --json gpurun_out/s1_scatter_ab.json > gpurun_out/s1_scatter_ab.log 2>&1; echo "scatter rc=$?"; tail -8 gpurun_out/s1_scatter_ab.log
timeout 300 python bench.py > gpurun_out/s1_bench_n1.json 2> gpurun_out/s1_bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/s1_bench_n1.json
