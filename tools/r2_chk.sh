set -u
mkdir -p gpurun_out
PYT="python -u -m pytest -q -m gpu --timeout=600 --timeout-method=thread -p no:cacheprovider"
run() { local secs=$1 name=$2; shift 2; timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
run 600 k_score $PYT tests/test_gpu_score.py
run 600 k_vit $PYT tests/test_gpu_vit.py tests/test_gpu_gemm.py
run 600 k_bench python bench.py --steps 20 --warmup 5 --no-cpu
tail -1 gpurun_out/k_bench.log > gpurun_out/k_bench.json
run 600 k_c4_p32 python tools/bench_configs.py --config c4 --patch 32 --steps 5 --no-cpu
run 600 k_c4_p16 python tools/bench_configs.py --config c4 --patch 16 --steps 5 --no-cpu
ncu --set full --clock-control none -k regex:gather_rows -s 6 -c 1 -o gpurun_out/gather python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > gpurun_out/k_ncu_gather.log 2>&1
ncu -i gpurun_out/gather.ncu-rep --page raw --csv > gpurun_out/gather.raw.csv 2>/dev/null
python tools/ncu_traffic.py gpurun_out/gather.raw.csv 4096 > gpurun_out/k_traffic.log 2>&1; cp profiles/gather_traffic.json gpurun_out/ 2>/dev/null; tail -1 gpurun_out/k_traffic.log
