"""Collects what tools/round2_gpu.sh left under gpurun_out/ into one markdown page (copy it to profiles/ once reviewed).
    python tools/summarize_round2.py > profiles/r02_staged_validation.md"""
import glob
import json
import os
import re

OUT = "gpurun_out"


def last_line(path):
    try:
        lines = [x for x in open(path, errors="replace").read().splitlines() if x.strip()]
        return lines[-1] if lines else ""
    except OSError:
        return ""


print("# Staged kernels on a B200: parity and timing (tools/round2_gpu.sh)\n")
print("## pytest stages\n\n| stage | result |\n|---|---|")
for log in sorted(glob.glob(os.path.join(OUT, "r2_*.log"))):
    tail = last_line(log)
    if re.search(r"\b(passed|failed|error|skipped)\b", tail):
        print(f"| `{os.path.basename(log)[:-4]}` | {tail.strip('= ')} |")
print("\n## bench lines\n\n| run | value | ms/step | e2e | note |\n|---|---|---|---|---|")
for js in sorted(glob.glob(os.path.join(OUT, "r2_*.json"))):
    try:
        d = json.loads(open(js).read().strip().splitlines()[-1])
    except Exception:
        continue
    if "value" in d:
        e2e = d.get("e2e", {}).get("value")
        sc = d.get("roofline_score_topk", {})
        note = f"score_topk {sc.get('achieved', 0):.0f} TFLOP/s" if sc.get("achieved") else d.get("config", {}).get("parallelism", "")
        print(f"| `{os.path.basename(js)[:-5]}` | {d['value']:.0f} {d.get('unit', '')} | {d.get('ms_per_step', 0):.3f} | "
              f"{(f'{e2e:.0f}' if e2e else '-')} | {note} |")
for name in ("bench_score.json", "r2_bench_score.json", "bench_attn_long.json"):
    p = os.path.join(OUT, name)
    if os.path.exists(p):
        print(f"\n## {name}\n\n```json\n{open(p).read().strip()}\n```")
