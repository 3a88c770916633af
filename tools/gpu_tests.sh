#!/bin/bash
# runs each GPU test file in its own process (a faulting kernel must not poison the others); logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rc=0
for f in tests/test_gpu_rows.py tests/test_gpu_ln_loss.py tests/test_gpu_attn.py tests/test_gpu_sasrec.py "$@"; do
  [ -f "$f" ] || continue
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -x --timeout=300 > gpurun_out/$n.log 2>&1
  r=$?
  echo "$n rc=$r $(tail -1 gpurun_out/$n.log)"
  [ $r -ne 0 ] && rc=1
done
exit $rc
