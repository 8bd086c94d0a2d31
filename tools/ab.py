#!/usr/bin/env python
"""Same-box A/B of library builds: alternates the candidates round-robin (each in its own process via
AURORA_B200_LIB) so box-to-box and thermal drift cancel.

  python tools/ab.py libA.so libB.so [...]  [--rows N] [--rounds R]
prints per library: median kernel_ms / total_ms over all rounds."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CHILD = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %r)
from aurora_b200.engine import DeviceBuffer, Index, to_bf16_bits
n = int(sys.argv[1]); d, nq, k = 768, 256, 32
rng = np.random.default_rng(1002)
block = to_bf16_bits(rng.standard_normal((50_000, d)).astype(np.float32))
q = to_bf16_bits(np.random.default_rng(2002).standard_normal((nq, d)).astype(np.float32))
with Index(d, n) as ix:
    for lo in range(0, n, 50_000):
        m = min(50_000, n - lo)
        ix.add(np.roll(block[:m], lo // 50_000, axis=1), np.arange(lo, lo + m, dtype=np.int64))
    dq = DeviceBuffer(q.nbytes).upload(q); ds = DeviceBuffer(nq * k * 4); di = DeviceBuffer(nq * k * 8)
    for _ in range(100):
        ix.search_dev(dq.ptr, nq, k, ds.ptr, di.ptr)
    ix.sync()
    ks, ts = [], []
    for _ in range(40):
        for _ in range(5):
            ix.search_dev(dq.ptr, nq, k, ds.ptr, di.ptr)
        ix.sync(); st = ix.stats(); ks.append(st["last_kernel_ms"]); ts.append(st["last_total_ms"])
    ids = di.download(np.empty((nq, k), dtype=np.int64))
print(json.dumps({"kernel_ms": float(np.median(ks)), "total_ms": float(np.median(ts)), "ids_sum": int(ids.sum())}))
''' % ROOT


def main():
    args = sys.argv[1:]
    rows, rounds = 1_000_000, 3
    libs = []
    while args:
        a = args.pop(0)
        if a == "--rows": rows = int(args.pop(0))
        elif a == "--rounds": rounds = int(args.pop(0))
        else: libs.append(a)
    res = {l: [] for l in libs}
    for r in range(rounds):
        for l in libs:
            env = dict(os.environ, AURORA_B200_AB_OLD_ABI="1", AURORA_B200_LIB=os.path.join(ROOT, "aurora_b200", l) if not os.path.isabs(l) else l)
            p = subprocess.run([sys.executable, "-c", CHILD, str(rows)], capture_output=True, text=True, env=env)
            if p.returncode != 0:
                print(l, "FAILED", p.stderr[-500:]); continue
            d = json.loads(p.stdout.strip().splitlines()[-1]); res[l].append(d)
            print(f"round {r} {l:28s} kernel {d['kernel_ms']:.4f} total {d['total_ms']:.4f} ids_sum {d['ids_sum']}", flush=True)
    import numpy as np
    for l in libs:
        if res[l]:
            print(f"== {l:28s} kernel {np.median([x['kernel_ms'] for x in res[l]]):.4f}  total {np.median([x['total_ms'] for x in res[l]]):.4f}  rows {rows}")


if __name__ == "__main__":
    main()
