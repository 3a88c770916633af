set -u
mkdir -p gpurun_out
timeout -s KILL 120 python tools/gemm_clocks.py cublas 2>&1 | tail -1 | tee -a gpurun_out/g4_clocks.jsonl
timeout -s KILL 120 python tools/gemm_clocks.py ours 2>&1 | tail -1 | tee -a gpurun_out/g4_clocks.jsonl
timeout -s KILL 120 env PR_GEMM_DEBUG=6 python tools/gemm_clocks.py ours 2>&1 | tail -1 | tee -a gpurun_out/g4_clocks.jsonl
timeout -s KILL 120 env PR_GEMM_DEBUG=2 python tools/gemm_clocks.py ours 2>&1 | tail -1 | tee -a gpurun_out/g4_clocks.jsonl
timeout -s KILL 120 env PR_GEMM_CG=1 python tools/gemm_clocks.py ours 2>&1 | tail -1 | tee -a gpurun_out/g4_clocks.jsonl
