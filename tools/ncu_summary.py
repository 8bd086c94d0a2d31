#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU): headline metrics + hottest SASS lines with stall reasons.
usage: python tools/ncu_summary.py <report.ncu-rep> [n_top] [kernel-name regex]"""
import csv, io, subprocess, sys

KSEL = []

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + KSEL, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (u, v) for h, u, v in zip(hdr, units, vals)}

def source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + KSEL, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if "Address" in r][0]
    return rows[0], rows[hi], rows[hi + 1:]

def main():
    rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    if len(sys.argv) > 3:
        KSEL.extend(["--kernel-name", "regex:" + sys.argv[3]])
    m = raw(rep)
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
            "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "smsp__inst_executed.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "launch__grid_size",
            "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    for k in keys:
        if k in m: print(f"{k:70s} {m[k][1]} {m[k][0]}")
    title, hdr, data = source(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    def f(r, h):
        try: return float(r[ix[h]])
        except Exception: return 0.0
    tot = sum(f(r, "# Samples") for r in data) or 1.0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(f(r, h) for r in data) for h in stalls}
    print("samples", int(tot), "sass lines", len(data))
    print("stall mix:", ", ".join(f"{k[6:]}={100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for i in sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:ntop]:
        r = data[i]
        dom = sorted(((h[6:], int(f(r, h))) for h in stalls), key=lambda x: -x[1])[:2]
        print(f"{i:5d} {r[ix['Address']][-5:]} {int(f(r,'# Samples')):6d} {100*f(r,'# Samples')/tot:5.1f}% x{int(f(r,'Instructions Executed')):<9d} {r[ix['Source']].strip()[:70]:70s} {[d for d in dom if d[1]>0]}")

if __name__ == "__main__":
    main()
