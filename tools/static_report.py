"""Static evidence per kernel: ptxas resource usage and the SASS mnemonics that show which hardware paths it uses
(tcgen05 = UTCHMMA[.2CTA] / LDTM / UTCBAR, TMA = UTMALDG[.MULTICAST] / UTMASTG / UBLKCP, mma.sync = HMMA-class MMA ops).
    python tools/static_report.py --all > profiles/r02k_sass_census.md     (every kernel of the library; needs nvcc + cuobjdump, no GPU)
    python tools/static_report.py > ...                                      (round 1's subset: the then-staged kernels)"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pixelrec_b200", "csrc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]
ALL = "--all" in sys.argv
FILES = {"score.cu": r"score_topk2_kernel|score_ce_merge|score_to_f16", "attn_long.cu": r"attn_long", "peer.cu": r"peers_kernel"}
MNEMONICS = ["UTCHMMA", "LDTM", "UTCBAR.MULTICAST", "UTCBAR", "UTMALDG.2D.MULTICAST", "UTMALDG.2D", "UCGABAR_ARV", "HMMA", "MUFU.EX2",
             "LDG.E.NA.128", "STG.E.128", "ATOMG"]
if ALL:
    FILES = {f: r"." for f in sorted(os.listdir(CSRC)) if f.endswith(".cu") and f != "api.cu"}
    MNEMONICS = ["UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "MUFU.EX2", "LDG.E.NA.128", "LDS.128",
                 "STG.E.128", "ATOMG", "SHFL"]
    FLAGS = FLAGS[:-2] + ["-DPR_SEED_DEV", "-Xptxas", "-v"]


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]


print("# r02k -- SASS census of every kernel in libpixelrec_b200.so (ptxas -v + cuobjdump -sass, sm_100a, the default build's flags)\n"
      if ALL else "# r01n -- static report of the staged kernels (ptxas -v + cuobjdump -sass, sm_100a; no GPU involved)\n")
if ALL:
    print("UTCHMMA = tcgen05.mma (count includes the .2CTA form of the CTA-pair GEMM), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, "
          "UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, HMMA = mma.sync.\n")
print("Counts are static instruction counts in the SASS of each kernel (loops not weighted).\n")
tmp = tempfile.mkdtemp()
for src, pat in FILES.items():
    obj = os.path.join(tmp, src + ".o")
    r = subprocess.run(["nvcc"] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stderr)
    res = {}
    cur = None
    for line in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", line)
        if m and cur:
            res.setdefault(cur, {})["stack"] = (int(m.group(1)), int(m.group(2)))
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            res.setdefault(cur, {})["regs"] = int(m.group(1))
    print(f"## {src}\n\n| kernel | regs | stack / spill B | SASS instr | " + " | ".join(MNEMONICS) + " |")
    print("|---|---:|---:|---:|" + "---:|" * len(MNEMONICS))
    for fn in sorted(res):
        name = demangle(fn)
        if not re.search(pat, name):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True).stdout
        ins = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
        counts = []
        for mn in MNEMONICS:
            n = sum(1 for l in ins if re.search(r"\b" + re.escape(mn) + r"(\b|\.)", l))
            if mn in ("UTCBAR", "UTMALDG.2D") and not ALL:      # plain forms only
                n -= sum(1 for l in ins if mn + ".MULTICAST" in l)
            counts.append(n)
        short = re.sub(r"^(void )?pr::", "", name)
        print(f"| `{short}` | {res[fn].get('regs', '?')} | {res[fn]['stack'][0]} / {res[fn]['stack'][1]} | {len(ins)} | "
              + " | ".join(str(c) if c else "" for c in counts) + " |")
    print()
