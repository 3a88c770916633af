mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_rows.py tests/test_gpu_ln_loss.py tests/test_gpu_sasrec.py tests/test_gpu_e2e.py -m gpu -q --maxfail=20 > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s6_pytest.log | cut -c1-300
for v in 0 4; do PR_SCATTER_VARIANT=$v timeout 120 python tools/bench_scatter.py --json gpurun_out/s6_scatter_v$v.json > gpurun_out/s6_scatter_v$v.log 2>&1; echo "scatter v$v rc=$?"; cut -c1-420 gpurun_out/s6_scatter_v$v.log; done
timeout 300 python bench.py > gpurun_out/s6_bench_n1.json 2> gpurun_out/s6_bench_n1.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/s6_bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s6_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/s6_ncu_bench.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/s6_launches.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"scatter_add_rows_ring" -c 1 -o gpurun_out/s6_ring python bench.py --steps 1 --warmup 1 --no-cpu --no-graph > gpurun_out/s6_ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/s6_ring.ncu-rep
