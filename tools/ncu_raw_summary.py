"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of numbers DESIGN.md / profiles/ quote.
    python tools/ncu_raw_summary.py gpurun_out/x.raw.csv > profiles/rNN_x.md"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    for d in data:
        name = d[ix["Kernel Name"]]
        print(f"### `{name[:110]}`\n")
        print("| metric | value |\n|---|---|")
        for key, label in WANT:
            if key in ix and d[ix[key]]:
                print(f"| {label} | {d[ix[key]]} {units[ix[key]]} |")
        top = sorted(((float(d[ix[h]].replace(",", "") or 0), h) for h in stalls if d[ix[h]]), reverse=True)[:6]
        print("| top stalls (warps per issue) | " + ", ".join(
            f"{h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in top) + " |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
