mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/s2_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/s2_pytest_all.log
timeout 200 python tools/bench_scatter.py --json gpurun_out/s2_scatter_ab.json > gpurun_out/s2_scatter_ab.log 2>&1; echo "scatter rc=$?"; cut -c1-330 gpurun_out/s2_scatter_ab.log
timeout 300 python bench.py > gpurun_out/s2_bench_n1.json 2> gpurun_out/s2_bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/s2_bench_n1.json
PR_FUSE_LN_Z=0 timeout 300 python bench.py --no-cpu > gpurun_out/s2_bench_n1_noz.json 2> gpurun_out/s2_bench_n1_noz.err; echo "bench noz rc=$?"; cut -c1-300 gpurun_out/s2_bench_n1_noz.json
