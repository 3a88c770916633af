"""SM clock / power while one GEMM implementation runs back to back for ~1.5 s: separates 'bubbles in the pipeline' (same clock,
fewer FLOP per clock) from 'draws more power' (lower clock under the 1 kW cap).  python tools/gemm_clocks.py cublas|ours"""
import json, os, subprocess, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelrec_b200 import ops
impl = sys.argv[1] if len(sys.argv) > 1 else "ours"
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
M = 81920
rows = []
res = {}
for name, K, N in [("K512_N1536", 512, 1536), ("K1536_N512", 1536, 512)]:
    x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.02
    y = torch.empty(M, N, device=dev)
    fn = (lambda: torch.mm(x, W.t(), out=y)) if impl == "cublas" else (lambda: ops.gemm(x, W, out=y))
    for _ in range(5): fn()
    torch.cuda.synchronize()
    proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "20"],
                            stdout=subprocess.PIPE, text=True)
    samples = []
    th = threading.Thread(target=lambda: [samples.append(l) for l in proc.stdout], daemon=True); th.start()
    time.sleep(0.2)
    n = 0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 1.5:
        for _ in range(50): fn()
        n += 50
        torch.cuda.synchronize()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    time.sleep(0.05); proc.terminate()
    vals = [l.strip().split(",") for l in samples if "," in l]
    clk = [float(v[0]) for v in vals[5:]]; pw = [float(v[1]) for v in vals[5:]]
    tf = 2.0 * M * N * K / ms / 1e9
    mhz = float(np.median(clk)) if clk else float("nan")
    res[name] = dict(ms=ms, TFLOPs=tf, sm_mhz_median=mhz, power_w_median=float(np.median(pw)) if pw else None,
                     flop_per_clk_per_sm=tf * 1e12 / (mhz * 1e6) / 148 if clk else None, n=len(clk))
print(json.dumps(dict(impl=impl, debug=os.environ.get("PR_GEMM_DEBUG", "0"), **res)))
