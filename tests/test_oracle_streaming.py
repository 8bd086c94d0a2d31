"""The streaming / threaded checker bench.py uses at full size must give the pinned oracle's answer."""

import numpy as np

from oracle import cosine_topk as O
from oracle.streaming_topk import FlatIndexF32, StreamingTopk, cosine_topk_streaming


def _data(n, d, nq, seed):
    rng = np.random.default_rng(seed)
    C = O.round_to_bf16(rng.standard_normal((n, d)).astype(np.float32))
    Q = O.round_to_bf16(rng.standard_normal((nq, d)).astype(np.float32))
    C[n // 2] = C[3]                      # exact duplicate: tie broken by id
    C[7] = 0.0                            # zero row: cosine 0.0
    Q[1] = C[3]
    return C, Q


def test_streaming_equals_pinned_oracle():
    for n, d, nq, k, chunk in [(5000, 96, 17, 10, 700), (20000, 768, 9, 32, 4096), (50, 64, 3, 8, 16), (300, 32, 4, 100, 64)]:
        C, Q = _data(n, d, nq, seed=n + k)
        ext = np.arange(n, dtype=np.int64) * 5 + 2
        gi, gs = cosine_topk_streaming(Q, C, k, ids=ext, chunk=chunk)
        oi, osc = O.cosine_topk(Q, C, k, ids=ext)
        assert np.array_equal(gi, oi)
        assert np.array_equal(gs, osc)       # same exact_cosine re-score -> identical floats


def test_streaming_chunking_does_not_matter_and_empty():
    C, Q = _data(3000, 64, 5, seed=1)
    a = cosine_topk_streaming(Q, C, 12, chunk=100)
    b = cosine_topk_streaming(Q, C, 12, chunk=3000)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    st = StreamingTopk(Q, 4)
    ids, sc = st.finish()
    assert (ids == -1).all() and np.isneginf(sc).all()


def test_flat_index_f32_agrees_with_oracle_up_to_fp32():
    C, Q = _data(8000, 128, 6, seed=3)
    ix = FlatIndexF32(128)
    ix.add(C[:5000]); ix.add(C[5000:])
    ids, sc = ix.search(Q, 8)
    oi, osc = O.cosine_topk(Q, C, 8)
    assert np.abs(sc - osc).max() < 1e-5
    assert (ids[:, 0] == oi[:, 0]).all() or np.abs(sc[:, 0] - osc[:, 0]).max() < 1e-6
    assert ix.rows == 8000 and ix.threads >= 1
