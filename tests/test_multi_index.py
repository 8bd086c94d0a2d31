"""MultiIndex (one owner process, one shard per GPU) against a single shard, on CPU doubles: the host-side logic --
id-mod-n placement, upserts / deletes reaching the owning shard, concurrent per-shard search, the (score desc, id asc)
merge, subset pre-filters, snapshots that restore onto a different number of shards, the retriever on top."""

import numpy as np
import pytest

from aurora_b200 import retriever as R
from aurora_b200.engine import MultiIndex
from oracle import cosine_topk as O
from tests.doubles import HashEmbedder, OracleIndex


def _mk(n_shards, dim=48, cap=4096):
    return MultiIndex(dim, cap, devices=list(range(n_shards)), shard_factory=lambda d, c, dev: OracleIndex(d, c))


def _same(a, b):
    assert np.array_equal(a[0], b[0])
    fin = np.isfinite(b[1])
    assert np.array_equal(np.isfinite(a[1]), fin) and np.allclose(a[1][fin], b[1][fin], atol=1e-6)


@pytest.mark.parametrize("n_shards", [1, 3, 8])
def test_matches_a_single_shard(n_shards):
    rng = np.random.default_rng(n_shards)
    n, d, nq, k = 1500, 48, 9, 12
    C = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    ids = rng.permutation(10 * n)[:n].astype(np.int64)
    ru, ro = rng.integers(0, 4, n).astype(np.int32), rng.integers(-1, 2, n).astype(np.int32)
    qu, qo = rng.integers(0, 4, nq).astype(np.int32), rng.integers(-1, 2, nq).astype(np.int32)
    one, many = OracleIndex(d, 4096), _mk(n_shards)
    for ix in (one, many):
        ix.add(C[:1000], ids[:1000], ru[:1000], ro[:1000])
        ix.add(C[1000:], ids[1000:], ru[1000:], ro[1000:])
        assert ix.remove(ids[5:300:7]) == len(ids[5:300:7])
        ix.add(C[:40] * 0.5 + 1.0, ids[100:140], ru[100:140], ro[100:140])      # upserts: the old rows must disappear
    assert many.stats()["rows"] == one.stats()["rows"] and many.stats()["live"] == one.stats()["live"]
    assert sum(many.stats()["rows_per_shard"]) == one.stats()["rows"] and len(many.stats()["rows_per_shard"]) == n_shards
    _same(many.search(Q, k), one.search(Q, k))
    _same(many.search(Q, k, qu, qo), one.search(Q, k, qu, qo))
    allow = ids[rng.permutation(n)[:200]]
    _same(many.search_subset(Q, k, allow), one.search_subset(Q, k, allow))
    _same(many.search(Q, 2000)[0:2], one.search(Q, 2000)[0:2])                  # k beyond the live rows: -1 / -inf padding last
    assert many.compact() == one.compact()
    _same(many.search(Q, k), one.search(Q, k))
    many.close()


def test_ties_break_by_ascending_id_across_shards():
    d = 16
    v = np.ones((1, d), dtype=np.float32)
    mi = _mk(4, dim=d)
    ids = np.array([11, 4, 9, 6, 2, 7], dtype=np.int64)                         # the same vector everywhere: all scores tie
    mi.add(np.repeat(v, len(ids), axis=0), ids)
    got, sc = mi.search(v, 4)
    assert got[0].tolist() == [2, 4, 6, 7] and np.allclose(sc, 1.0)
    mi.close()


def test_snapshot_restores_onto_another_shard_count(tmp_path):
    rng = np.random.default_rng(3)
    C = O.round_to_bf16(rng.standard_normal((400, 48)).astype(np.float32))
    Q = rng.standard_normal((5, 48)).astype(np.float32)
    a = _mk(3)
    a.add(C, np.arange(400, dtype=np.int64), np.zeros(400, np.int32), np.full(400, -1, np.int32))
    a.remove(np.arange(0, 400, 9))
    want = a.search(Q, 7)
    a.save(str(tmp_path / "shard"))
    b = MultiIndex.load(str(tmp_path / "shard"), devices=[0, 1, 2, 3, 4], shard_factory=lambda d, c, dev: OracleIndex(d, c))
    _same(b.search(Q, 7), want)
    assert b.stats()["live"] == a.stats()["live"] == b.stats()["rows"]           # tombstones are not saved
    a.close(); b.close()


def test_retriever_on_top_of_a_multi_index():
    emb = HashEmbedder(64)
    kb1 = R.KnowledgeBase(emb, capacity=512, index_factory=lambda dim, cap: OracleIndex(dim, cap))
    kb8 = R.KnowledgeBase(emb, capacity=512, index_factory=lambda dim, cap: MultiIndex(
        dim, cap, devices=list(range(8)), shard_factory=lambda d, c, dev: OracleIndex(d, c)))
    docs = [f"runbook {i}: restart service svc{i % 7} after alert code zx{i}" for i in range(60)]
    for kb in (kb1, kb8):
        for t in range(3):
            kb.insert(f"user{t}", f"doc{t}", "r.md", [{"content": c, "chunk_index": i} for i, c in enumerate(docs[t * 20:(t + 1) * 20])],
                      org_id="org" if t < 2 else None)
        kb.delete_where(lambda p: p["document_id"] == "doc1" and p["chunk_index"] % 3 == 0)
    for q, u in (("restart service svc3", "user0"), ("alert code zx41", "user2"), ("runbook", "user1")):
        a = [(o.properties["document_id"], o.properties["chunk_index"], round(o.metadata.score, 6)) for o in kb1.query(q, 6, user_id=u)]
        b = [(o.properties["document_id"], o.properties["chunk_index"], round(o.metadata.score, 6)) for o in kb8.query(q, 6, user_id=u)]
        assert a == b and a


def test_native_host_merge_matches_numpy_order():
    """csrc/host_merge.cpp (the owner process's k-way merge; pure host code, callable without a GPU) against the
    numpy statement of the same order, on sorted lists with ties and padded tails."""
    import ctypes as C

    from aurora_b200 import _native as N

    lib = C.CDLL(N.LIB_PATH)
    lib.aur_merge_topk_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.aur_merge_topk_host.restype = C.c_int
    vp = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    rng = np.random.default_rng(0)
    for n_lists, nq, k_in, k_out in ((3, 7, 5, 5), (8, 50, 32, 32), (2, 4, 6, 9), (64, 3, 4, 10)):
        sc = np.round(rng.standard_normal((n_lists, nq, k_in)), 1).astype(np.float32)      # one decimal: plenty of ties
        ids = rng.permutation(n_lists * nq * k_in).reshape(n_lists, nq, k_in).astype(np.int64)
        for l in range(n_lists):
            for q in range(nq):
                o = np.lexsort((ids[l, q], -sc[l, q].astype(np.float64)))
                sc[l, q], ids[l, q] = sc[l, q][o], ids[l, q][o]
                if rng.random() < 0.3:
                    cut = rng.integers(0, k_in + 1)
                    ids[l, q, cut:], sc[l, q, cut:] = -1, -np.inf
        out_s, out_i = np.empty((nq, k_out), np.float32), np.empty((nq, k_out), np.int64)
        assert lib.aur_merge_topk_host(vp(sc), vp(ids), n_lists, nq, k_in, k_out, vp(out_s), vp(out_i)) == 0
        I, S = np.concatenate(list(ids), axis=1), np.concatenate(list(sc), axis=1)
        order = np.lexsort((np.where(I < 0, np.iinfo(np.int64).max, I), -S.astype(np.float64)), axis=1)
        wi, ws = np.take_along_axis(I, order, axis=1), np.take_along_axis(S, order, axis=1)
        if k_out > wi.shape[1]:
            wi = np.concatenate([wi, np.full((nq, k_out - wi.shape[1]), -1)], axis=1)
            ws = np.concatenate([ws, np.full((nq, k_out - ws.shape[1]), -np.inf, np.float32)], axis=1)
        assert np.array_equal(out_i, wi[:, :k_out]) and np.array_equal(out_s, ws[:, :k_out])
    assert lib.aur_merge_topk_host(None, None, 1, 1, 1, 1, None, None) != 0
