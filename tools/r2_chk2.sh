set -u
mkdir -p gpurun_out
PYT="python -u -m pytest -q -m gpu --timeout=600 --timeout-method=thread -p no:cacheprovider"
run() { local secs=$1 name=$2; shift 2; timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
run 600 m_attn_vit $PYT tests/test_gpu_attn_long.py tests/test_gpu_vit.py tests/test_gpu_attn.py
for c in "c4 --patch 32 --steps 5" "c4 --patch 16 --steps 5" "c5" "c3"; do
  n=$(echo $c | tr -d ' -' ); run 600 m_$n python tools/bench_configs.py --config $c --no-cpu
  tail -1 gpurun_out/m_$n.log > gpurun_out/m_$n.json
done
run 600 m_c4_p32_eager python tools/bench_configs.py --config c4 --patch 32 --steps 5 --no-cpu --no-graph
