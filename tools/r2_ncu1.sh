set -u
mkdir -p gpurun_out
export PR_EXPERIMENTAL=1
NCU="ncu --set full --clock-control none --import-source on"
timeout -s KILL 300 python tools/bench_linear.py > gpurun_out/r2_bench_linear.log 2>&1; echo "bench_linear rc=$?"; tail -2 gpurun_out/r2_bench_linear.log | cut -c1-600
timeout -s KILL 400 env SCORE_TUNES=16 $NCU -k regex:score_topk2 -c 1 -o gpurun_out/r2_score_v2 python tools/bench_score.py > gpurun_out/r2_ncu_score_v2.log 2>&1; echo "ncu v2 rc=$?"
timeout -s KILL 400 env PR_TUNE=9 SCORE_F16=1 $NCU -k regex:score_topk2 -c 1 -o gpurun_out/r2_score_f16 python tools/bench_score.py > gpurun_out/r2_ncu_score_f16.log 2>&1; echo "ncu f16 rc=$?"
for f in gpurun_out/r2_*.ncu-rep; do [ -f "$f" ] && ncu -i "$f" --page raw --csv > "${f%.ncu-rep}.raw.csv" 2>/dev/null; done
ls -la gpurun_out | head -50
