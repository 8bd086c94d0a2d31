#!/usr/bin/env python
"""simtopk_tc_kernel time against shard size on ONE GPU (same clocks for every point): the slope is the
streaming rate, the intercept the non-streaming term (launch, query load, bootstrap, tail).

  python tools/rows_sweep.py [out.json] [nq] [k]      (numpy + ctypes only)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aurora_b200.engine import DeviceBuffer, Index, to_bf16_bits  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/simtopk_rows_sweep.json"
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    d = int(os.environ.get("AUR_DIM", "768"))
    sizes = [62_500, 125_000, 250_000, 500_000, 1_000_000, 2_000_000, 4_000_000]
    nmax = max(sizes)
    rng = np.random.default_rng(1002)
    block = to_bf16_bits(rng.standard_normal((50_000, d)).astype(np.float32))
    q = to_bf16_bits(np.random.default_rng(2002).standard_normal((nq, d)).astype(np.float32))
    pts = []
    dq = DeviceBuffer(q.nbytes).upload(q)
    ds = DeviceBuffer(nq * k * 4)
    di = DeviceBuffer(nq * k * 8)
    with Index(d, nmax) as ix:
        have = 0
        for n in sizes:
            while have < n:
                m = min(50_000, n - have)
                ix.add(np.roll(block[:m], have // 50_000, axis=1), np.arange(have, have + m, dtype=np.int64))
                have += m
            ks, ts = [], []
            for rep in range(24):
                ix.search_dev(dq.ptr, nq, k, ds.ptr, di.ptr)
                ix.sync()
                st = ix.stats()
                if rep >= 4:
                    ks.append(st["last_kernel_ms"]); ts.append(st["last_total_ms"])
            pts.append({"rows": n, "kernel_ms": float(np.median(ks)), "kernel_ms_min": float(np.min(ks)),
                        "total_ms": float(np.median(ts)), "bytes": n * d * 2})
            print(pts[-1], flush=True)
    x = np.array([p["rows"] for p in pts if p["rows"] >= 500_000], dtype=np.float64)
    y = np.array([p["kernel_ms"] for p in pts if p["rows"] >= 500_000])
    slope, icpt = np.polyfit(x, y, 1)
    res = {"nq": nq, "k": k, "dim": d, "points": pts, "fit_rows_ge_500k": {"ms_per_Mrow": slope * 1e6, "intercept_ms": icpt,
           "stream_GBps": d * 2 / (slope * 1e-3) / 1e9}}
    print(json.dumps(res["fit_rows_ge_500k"]))
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
