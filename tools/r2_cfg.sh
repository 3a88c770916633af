set -u
mkdir -p gpurun_out
PYT="python -u -m pytest -q -m gpu --timeout=600 --timeout-method=thread -p no:cacheprovider"
run() { local secs=$1 name=$2; shift 2; timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-500)"; return $rc; }
run 300 c_test_gemm $PYT tests/test_gpu_gemm.py
run 600 c_sasrec $PYT tests/test_gpu_sasrec.py tests/test_gpu_e2e.py
for c in c5 c3; do
  run 600 cfg_$c python tools/bench_configs.py --config $c
  tail -1 gpurun_out/cfg_$c.log > gpurun_out/cfg_$c.json
done
run 600 cfg_c4_p32 python tools/bench_configs.py --config c4 --patch 32 --steps 5
tail -1 gpurun_out/cfg_c4_p32.log > gpurun_out/cfg_c4_p32.json
run 600 cfg_c4_p16 python tools/bench_configs.py --config c4 --patch 16 --steps 5
tail -1 gpurun_out/cfg_c4_p16.log > gpurun_out/cfg_c4_p16.json
