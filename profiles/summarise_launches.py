"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --clock-control none
--csv --log-file X.csv python bench.py ...`) per kernel: launches, total time, share of the summed kernel time, DRAM bytes.
Writes <out>.csv and, with --traffic-json, the per-launch DRAM traffic of every kernel (bench.py reads it for `roofline.traffic`).

  python profiles/summarise_launches.py gpurun_out/launches_bench.csv profiles/r1/launches_bench_by_kernel.csv --traffic-json profiles/r1/traffic.json
"""
import collections
import csv
import json
import sys

src, out = sys.argv[1], sys.argv[2]
tj = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
data = rows[rows.index(hdr) + 1:]
ik, im, iv, iu, iid = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
SCALE = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.defaultdict(lambda: {"ids": set(), "us": 0.0, "rd": 0.0, "wr": 0.0})


def short(name):
    n = name.split("(")[0].strip()
    for pre in ("void ", "avi::"):
        if n.startswith(pre):
            n = n[len(pre):]
    return n.replace("avi::", "")[:64]


for r in data:
    a = agg[short(r[ik])]
    a["ids"].add(r[iid])
    v = float(r[iv].replace(",", "")) * SCALE.get(r[iu], 1.0)
    if r[im].startswith("gpu__time_duration"):
        a["us"] += v
    elif r[im].startswith("dram__bytes_read"):
        a["rd"] += v
    elif r[im].startswith("dram__bytes_write"):
        a["wr"] += v
tot = sum(a["us"] for a in agg.values())
with open(out, "w") as fh:
    fh.write("kernel,launches,time_us,share,dram_read_MB,dram_write_MB\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        fh.write(f"{k},{len(a['ids'])},{a['us']:.1f},{a['us'] / tot:.3f},{a['rd'] / 1e6:.1f},{a['wr'] / 1e6:.1f}\n")
    fh.write(f"TOTAL,{sum(len(a['ids']) for a in agg.values())},{tot:.1f},1.000,{sum(a['rd'] for a in agg.values()) / 1e6:.1f},"
             f"{sum(a['wr'] for a in agg.values()) / 1e6:.1f}\n")
if tj:
    json.dump({k: {"launches": len(a["ids"]), "dram_bytes_per_launch": (a["rd"] + a["wr"]) / max(len(a["ids"]), 1),
                   "time_share": a["us"] / tot} for k, a in agg.items() if a["rd"] + a["wr"] > 0},
              open(tj, "w"), indent=1, sort_keys=True)
print(open(out).read())
