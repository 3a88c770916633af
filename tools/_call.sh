mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rows.py tests/test_gpu_sasrec.py tests/test_gpu_e2e.py tests/test_gpu_peer.py -m gpu -q --maxfail=20 > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s7_pytest.log | cut -c1-300
PR_SCATTER_VARIANT=8 timeout 200 python -m pytest tests/test_gpu_rows.py -m gpu -q --maxfail=20 > gpurun_out/s7_pytest_v1.log 2>&1; echo "pytest v1 rc=$?"; tail -2 gpurun_out/s7_pytest_v1.log | cut -c1-300
for v in 0 4 8; do PR_SCATTER_VARIANT=$v timeout 120 python tools/bench_scatter.py --quick --json gpurun_out/s7_scatter_v$v.json > gpurun_out/s7_scatter_v$v.log 2>&1; echo "scatter v$v rc=$?"; cut -c1-420 gpurun_out/s7_scatter_v$v.log; done
timeout 300 python bench.py > gpurun_out/s7_bench_n1.json 2> gpurun_out/s7_bench_n1.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/s7_bench_n1.json
