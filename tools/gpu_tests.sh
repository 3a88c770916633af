#!/bin/bash
# runs each GPU test file in its own process (a faulting kernel must not poison the others); logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rc=0
for f in tests/test_gpu_rows.py tests/test_gpu_ln_loss.py tests/test_gpu_attn.py tests/test_gpu_sasrec.py tests/test_gpu_e2e.py tests/test_gpu_dist.py tests/test_gpu_score.py "$@"; do
  [ -f "$f" ] || continue
  n=$(basename $f .py)
  timeout 900 python -u -m pytest $f -v -m gpu ${PYTEST_X:-} --timeout=120 --timeout-method=thread -p no:cacheprovider > gpurun_out/$n.log 2>&1
  r=$?
  echo "$n rc=$r $(tail -1 gpurun_out/$n.log)"; grep -E "FAILED|Timeout|rror" gpurun_out/$n.log | head -8
  [ $r -ne 0 ] && rc=1
done
exit $rc
