"""Property tests (hypothesis) of the host-side text machinery: BM25 index invariants, ranked fusion,
WordPiece robustness on arbitrary unicode, snapshot of the keyword index under upserts / deletes."""

import math

import pytest

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from aurora_b200.bm25 import BM25Index, ranked_fusion, tokenize  # noqa: E402
from aurora_b200.wordpiece import WordPieceTokenizer, basic_tokenize  # noqa: E402

words = st.sampled_from(["redis", "kafka", "lag", "cpu", "disk", "pod", "alert", "node", "zx9981", "latency", "p99"])
docs = st.lists(st.lists(words, min_size=1, max_size=12).map(" ".join), min_size=1, max_size=12)


@settings(max_examples=60, deadline=None)
@given(docs, st.lists(words, min_size=1, max_size=4).map(" ".join))
def test_bm25_scores_sorted_positive_and_only_matching_docs(texts, query):
    ix = BM25Index()
    for i, t in enumerate(texts):
        ix.add(i, t)
    res = ix.search(query, 100)
    qt = set(tokenize(query))
    assert all(s > 0 and math.isfinite(s) for _, s in res)
    assert [s for _, s in res] == sorted((s for _, s in res), reverse=True)
    assert {d for d, _ in res} == {i for i, t in enumerate(texts) if qt & set(tokenize(t))}
    assert ix.search(query, 3) == res[:3]                                  # limit is a prefix


@settings(max_examples=40, deadline=None)
@given(docs)
def test_bm25_upsert_and_remove_leave_no_trace(texts):
    a, b = BM25Index(), BM25Index()
    for i, t in enumerate(texts):
        a.add(i, t)
        b.add(i, "placeholder text that will be replaced")
        b.add(i, t)                                                        # upsert = same as fresh insert
    b.add(999, "temporary document about redis kafka cpu")
    assert b.remove(999)
    for q in ("redis", "kafka lag", "cpu disk pod", "placeholder", "temporary"):
        ra, rb = a.search(q, 50), b.search(q, 50)
        assert [d for d, _ in ra] == [d for d, _ in rb]
        assert all(abs(x - y) < 1e-12 for (_, x), (_, y) in zip(ra, rb))
    assert len(a) == len(b) == len(texts)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 30), unique=True, max_size=20), st.lists(st.integers(0, 30), unique=True, max_size=20),
       st.floats(0.0, 1.0), st.integers(1, 10))
def test_ranked_fusion_properties(dense, sparse, alpha, limit):
    fused = ranked_fusion([(alpha, dense), (1.0 - alpha, sparse)], limit)
    ids = [d for d, _ in fused]
    assert len(ids) == len(set(ids)) <= limit
    assert [s for _, s in fused] == sorted((s for _, s in fused), reverse=True)
    contributing = (set(dense) if alpha > 0 else set()) | (set(sparse) if alpha < 1 else set())
    assert set(ids) <= contributing
    for d, s in fused:                                                     # score = sum of weight / (rank + 60)
        want = (alpha / (dense.index(d) + 60) if d in dense and alpha > 0 else 0.0) + \
               ((1 - alpha) / (sparse.index(d) + 60) if d in sparse and alpha < 1 else 0.0)
        assert abs(s - want) < 1e-12
    if alpha == 1.0:
        assert ids == dense[:limit]


@settings(max_examples=80, deadline=None)
@given(st.text(max_size=200))
def test_wordpiece_never_crashes_and_respects_limits(text):
    vocab = {t: i for i, t in enumerate(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "a", "b", "##a", "##b", "the", "##s", "!", "."])}
    tok = WordPieceTokenizer(vocab)
    ids = tok.encode(text, max_len=16)
    assert 2 <= len(ids) <= 16 and ids[0] == vocab["[CLS]"] and ids[-1] == vocab["[SEP]"]
    assert all(0 <= i < len(vocab) for i in ids)
    assert all(w and not any(c.isspace() for c in w) for w in basic_tokenize(text))


# ---------------------------------------------------------------------- one owner, many shards (engine.MultiIndex on CPU doubles)
ops = st.lists(st.tuples(st.sampled_from(["add", "remove", "upsert"]), st.integers(0, 10 ** 6)), min_size=1, max_size=25)


@settings(max_examples=40, deadline=None)
@given(ops, st.integers(1, 6), st.integers(1, 9))
def test_multi_index_equals_one_shard_under_random_mutations(op_list, n_shards, k):
    import numpy as np

    from aurora_b200.engine import MultiIndex
    from tests.doubles import OracleIndex

    d = 16
    one = OracleIndex(d, 4096)
    many = MultiIndex(d, 4096, devices=list(range(n_shards)), shard_factory=lambda dd, c, dev: OracleIndex(dd, c))
    known = []
    for step, (op, seed) in enumerate(op_list):
        rng = np.random.default_rng(seed)
        if op == "add" or not known:
            ids = np.arange(len(known) * 7 + 1000 * step, len(known) * 7 + 1000 * step + int(rng.integers(1, 8)), dtype=np.int64)
            known += ids.tolist()
        elif op == "upsert":
            ids = np.array(sorted(set(rng.choice(known, size=min(len(known), 3)).tolist())), dtype=np.int64)
        else:
            gone = np.array(sorted(set(rng.choice(known, size=min(len(known), 3)).tolist())), dtype=np.int64)
            assert many.remove(gone) == one.remove(gone)
            continue
        rows = rng.standard_normal((len(ids), d)).astype(np.float32)
        users = rng.integers(0, 3, len(ids)).astype(np.int32)
        orgs = rng.integers(-1, 2, len(ids)).astype(np.int32)
        one.add(rows, ids, users, orgs); many.add(rows, ids, users, orgs)
    Q = np.random.default_rng(1).standard_normal((3, d)).astype(np.float32)
    for qu, qo in ((None, None), (np.array([0, 1, 2], np.int32), np.array([-1, 0, 1], np.int32))):
        a, b = many.search(Q, k, qu, qo), one.search(Q, k, qu, qo)
        assert np.array_equal(a[0], b[0])
        fin = np.isfinite(b[1])
        assert np.array_equal(np.isfinite(a[1]), fin) and np.allclose(a[1][fin], b[1][fin], atol=1e-6)
    assert many.stats()["live"] == one.stats()["live"]
    many.close()


@settings(max_examples=40, deadline=None)
@given(docs, st.lists(words, min_size=1, max_size=5).map(" ".join), st.integers(1, 8), st.booleans())
def test_bm25_array_path_equals_loop_path(texts, query, limit, scoped):
    """Posting lists past bm25.VECTORISE_FROM are scored as numpy arrays: same ranking and scores as the Python loop,
    with and without a tenant pre-filter, also when only some of the query's terms take the array path."""
    from aurora_b200 import bm25 as M

    ix = BM25Index()
    for rep in range(3):                                   # a few copies so that posting lists differ in length
        for i, t in enumerate(texts):
            if (i + rep) % 3 != 2:
                ix.add(rep * 100 + i, t)
    allowed = {d for d in ix._doc_len if d % 2 == 0} if scoped else None
    old = M.VECTORISE_FROM
    try:
        M.VECTORISE_FROM = 10 ** 9
        want = ix.search(query, limit, allowed=allowed)
        for thr in (1, 3, 6):
            M.VECTORISE_FROM = thr
            got = ix.search(query, limit, allowed=allowed)
            assert [d for d, _ in got] == [d for d, _ in want]
            assert all(abs(a - b) < 1e-9 for (_, a), (_, b) in zip(got, want))
    finally:
        M.VECTORISE_FROM = old
