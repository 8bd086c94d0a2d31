timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_multi.py -m gpu -x -q -k "multi" 2>&1 | tail -5
python - <<'PY' 2>&1 | tee gpurun_out/multi_index_latency_r2.txt
import time, numpy as np, sys
sys.path.insert(0, ".")
from aurora_b200 import _native as N
from aurora_b200.engine import Index, MultiIndex, to_bf16_bits
rng = np.random.default_rng(0)
n, d = 1000000, 768
nd = N.load().aur_device_count()
block = to_bf16_bits(rng.standard_normal((100000, d)).astype(np.float32))
print(f"# {n} x {d} bf16 rows; Index = one GPU; MultiIndex = one owner process, {nd} GPUs, host merge (aur_merge_topk_host); wall clock per call, host buffers")
for nq, k in ((1, 10), (64, 10), (256, 32)):
    Q = to_bf16_bits(rng.standard_normal((nq, d)).astype(np.float32))
    res = {}
    for name, mk in (("Index", lambda: Index(d, n)), (f"MultiIndex x{nd}", lambda: MultiIndex(d, n))):
        ix = mk()
        for lo in range(0, n, 100000):
            ix.add(np.roll(block, lo // 100000, axis=1), np.arange(lo, lo + 100000, dtype=np.int64))
        for _ in range(5): ix.search(Q, k)
        t0 = time.perf_counter()
        for _ in range(50): ix.search(Q, k)
        res[name] = (time.perf_counter() - t0) / 50 * 1e3
        ix.close()
    print(f"nq={nq} k={k}: " + ", ".join(f"{a} {b:.3f} ms" for a, b in res.items()), flush=True)
PY
