"""Helpers to read ncu exports (run in the build container: `ncu -i X.ncu-rep --page raw|source --csv`)."""
import collections
import csv
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def summary(rep):
    hdr, units, data = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print(d[idx["Kernel Name"]][:70])
        for k in KEYS:
            if k in idx:
                print(f"   {k:70s} {d[idx[k]]:>16s} {units[idx[k]]}")


def source(rep, launch=0, top=30):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # a report with several results prints one block per result ("Kernel Name" row, header row, SASS rows): keep the first block
    # (--launch-skip already selected the result) and drop incomplete rows
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    end = starts[1] if len(starts) > 1 else len(rows)
    rows = rows[starts[0]:end] if starts else rows
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("Warp Stall Sampling (All Samples)")] != ""]
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[i_s]) for r in data)
    print(rows[0][1][:80], "total samples", tot, "SASS instructions", len(data))
    agg = collections.Counter()
    for r in data:
        for i, h in stall:
            agg[h] += int(r[i] or 0)
    print("stall reasons:", [(h, n) for h, n in agg.most_common(8)])
    for r in sorted(data, key=lambda r: -int(r[i_s]))[:top]:
        reasons = sorted(((int(r[i] or 0), h) for i, h in stall), reverse=True)[:2]
        print(r[i_s].rjust(7), r[i_ex].rjust(10), r[i_src].strip()[:90].ljust(90), reasons)
    ops = collections.Counter()
    for r in data:
        t = r[i_src].strip().split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += int(r[i_ex])
    print("executed by opcode:", ops.most_common(22))


if __name__ == "__main__":
    if sys.argv[1] == "summary":
        summary(sys.argv[2])
    else:
        source(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
