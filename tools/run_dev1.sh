timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "multi_index" 2>&1 | tail -5
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, ".")
from aurora_b200.engine import Index, MultiIndex, to_bf16_bits
rng = np.random.default_rng(0)
n, d = 400000, 768
C = to_bf16_bits(rng.standard_normal((n, d)).astype(np.float32))
ids = np.arange(n, dtype=np.int64)
for nq, k in ((1, 10), (64, 10), (256, 32)):
    Q = to_bf16_bits(rng.standard_normal((nq, d)).astype(np.float32))
    res = {}
    for name, mk in (("Index", lambda: Index(d, n)), ("MultiIndex x3 (same GPU)", lambda: MultiIndex(d, n, devices=[0, 0, 0]))):
        ix = mk(); ix.add(C, ids)
        for _ in range(5): ix.search(Q, k)
        t0 = time.perf_counter()
        for _ in range(50): ix.search(Q, k)
        res[name] = (time.perf_counter() - t0) / 50 * 1e3
        ix.close()
    print(f"nq={nq} k={k}: " + ", ".join(f"{a} {b:.3f} ms" for a, b in res.items()), flush=True)
PY
