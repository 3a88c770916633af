set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-graph > gpurun_out/ll_bench.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/ll_summary.md 2>&1; head -50 gpurun_out/ll_summary.md
NCU="ncu --set full --clock-control none --import-source on"
timeout -s KILL 500 $NCU -k regex:gemm_tf32 -s 4 -c 3 -o gpurun_out/ll_gemm python tools/bench_linear.py --iters 1 > gpurun_out/ll_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ncu -i gpurun_out/ll_gemm.ncu-rep --page raw --csv > gpurun_out/ll_gemm.raw.csv 2>/dev/null
timeout -s KILL 300 $NCU -k regex:score_topk2 -c 1 -o gpurun_out/ll_score python tools/bench_score.py > gpurun_out/ll_ncu_score.log 2>&1
ncu -i gpurun_out/ll_score.ncu-rep --page raw --csv > gpurun_out/ll_score.raw.csv 2>/dev/null
