set -u
mkdir -p gpurun_out
run() { local secs=$1 name=$2; shift 2; timeout --signal=KILL "$secs" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
run 1200 full_pytest python -m pytest tests -x -q -m gpu
run 300 full_smoke python -c "import __graft_entry__ as g; g.smoke()"
run 600 full_bench python bench.py --steps 20 --warmup 5
tail -1 gpurun_out/full_bench.log > gpurun_out/full_bench.json
run 600 full_bench_ref python bench.py --impl reference --steps 3 --warmup 1
