"""profiles/gather_traffic.json from an `ncu --set full` capture of the gather kernel inside bench.py (B200_PROFILING.md):

    ncu --set full --clock-control none -k regex:gather_rows -s 6 -c 1 -o gpurun_out/gather python bench.py --steps 2 --warmup 3 --no-cpu --no-graph
    ncu -i gpurun_out/gather.ncu-rep --page raw --csv > gpurun_out/gather.raw.csv
    python tools/ncu_traffic.py gpurun_out/gather.raw.csv 4096

bench.py reports the number as roofline.traffic (per launch, like roofline.achieved)."""
import csv
import json
import os
import sys


def main(path, batch):
    rows = list(csv.reader(open(path)))
    hdr, units, d = rows[0], rows[1], rows[2]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(name):
        v, u = float(d[ix[name]].replace(",", "")), units[ix[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    out = {"kernel": d[ix["Kernel Name"]][:80], "batch": int(batch), "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_launch": rd + wr, "duration_us_under_ncu": float(d[ix["gpu__time_duration.sum"]].replace(",", "")),
           "source": os.path.basename(path)}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    json.dump(out, open(os.path.join(root, "profiles", "gather_traffic.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 4096)
