"""Tenant-scoped search timing (the reference's own call pattern: one scope per call).
   python tools/profile_tenant.py [n_rows]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aurora_b200 import _native as N
from aurora_b200.engine import Index, to_bf16_bits

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = 768
rng = np.random.default_rng(1)
blk = to_bf16_bits(rng.standard_normal((50_000, d)).astype(np.float32))
with Index(d, n) as ix:
    for lo in range(0, n, 50_000):
        m = min(50_000, n - lo)
        ix.add(np.roll(blk[:m], lo // 50_000, axis=1), np.arange(lo, lo + m, dtype=np.int64),
               rng.integers(0, 10, m).astype(np.int32), np.full(m, -1, np.int32))
    for nq, k in ((256, 32), (1, 5)):
        q = to_bf16_bits(rng.standard_normal((nq, d)).astype(np.float32))
        for scoped in (False, True, "mixed"):
            if scoped == "mixed" and nq == 1:
                continue
            qu = (rng.integers(0, 10, nq).astype(np.int32) if scoped == "mixed" else np.full(nq, 3, np.int32)) if scoped else None
            for _ in range(5):
                ix.search(q, k, qu, None)
            t0 = time.perf_counter()
            reps = 50
            for _ in range(reps):
                ix.search(q, k, qu, None)
            wall = (time.perf_counter() - t0) / reps * 1e3
            st = ix.stats()
            print(f"rows {n} nq {nq} k {k} scope {('10 different users in one batch (row bit masks)' if scoped == 'mixed' else 'user==3 (10% of rows)') if scoped else 'none'}: kernel {N.KERNEL_NAMES[st['last_kernel']]} "
                  f"device {st['last_total_ms']:.3f} ms, host call {wall:.3f} ms, launches {st['last_launches']}", flush=True)
