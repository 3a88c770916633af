"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step
(the launches between two consecutive table gathers) grouped by kernel.  Usage:
    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launch_list.md
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"<.*$", "", name) if not name.startswith("pixelrec") else name
    return name[:70]


def owner(name):
    if "pixelrec_b200" in name or re.search(r"\b(gather_rows|scatter_add|adamw_|add_ln|act_|attn_|bpr_|colsum|plan_|rs_|scan_|seg_|seq_batch|score_|mask_|gemm_tf32|gemm_splitk|rownorm)", name):
        return "ours"
    if "cutlass" in name or "gemm" in name.lower() or "cublas" in name.lower() or "splitK" in name:
        return "cuBLAS"
    if "nccl" in name.lower():
        return "NCCL"
    return "ATen"


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    head, rows = rows[0], rows[1:]
    k_name, k_val = head.index("Kernel Name"), head.index("Metric Value")
    launches = [(r[k_name], float(r[k_val].replace(",", "")) / 1e3) for r in rows]
    marks = [i for i, (n, _) in enumerate(launches) if "gather_rows" in n]
    if len(marks) < 2:
        raise SystemExit("need two table gathers to delimit a step")
    a, b = marks[-2], marks[-1]
    step = launches[a:b]
    total = sum(t for _, t in step)
    groups = OrderedDict()
    for n, t in step:
        g = groups.setdefault(short(n), [0, 0.0, owner(n)])
        g[0] += 1
        g[1] += t
    print(f"launches in the step: {len(step)}   sum of kernel durations: {total/1e3:.3f} ms (serialised, cold-cache)\n")
    by_owner = {}
    for _, (c, t, o) in groups.items():
        by_owner[o] = by_owner.get(o, 0.0) + t
    print("| owner | ms | share |\n|---|---|---|")
    for o, t in sorted(by_owner.items(), key=lambda kv: -kv[1]):
        print(f"| {o} | {t/1e3:.3f} | {100*t/total:.1f} % |")
    print("\n| kernel | owner | launches | total us | share |\n|---|---|---|---|---|")
    for n, (c, t, o) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {o} | {c} | {t:.1f} | {100*t/total:.1f} % |")


if __name__ == "__main__":
    main(sys.argv[1])
