"""The oracle is pinned here: against the vectors the reference's own tests pin
(server/tests/services/correlation/test_similarity_strategy.py:31-47, :124, :187, :193-216)
and against golden outputs produced by importing the real reference function
(tests/golden/cosine_ref.json, made by oracle/gen_golden.py).  CPU only."""

import json
import os

import numpy as np
import pytest

from oracle import cosine_topk as O
from oracle import ref_cosine as R

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "cosine_ref.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


# ---- known answers the reference's own tests assert -------------------------------------
def test_reference_known_answers():
    assert R.cosine_similarity([1, 2, 3], [1, 2, 3]) == pytest.approx(1.0)       # :193-196
    assert R.cosine_similarity([1, 0], [0, 1]) == pytest.approx(0.0)             # :198-201
    assert R.cosine_similarity([1, 0], [-1, 0]) == 0.0                           # :203-206 (clamped)
    assert R.cosine_similarity([], []) == 0.0                                    # :208-211
    assert R.cosine_similarity([1, 2], [1, 2, 3]) == 0.0                         # :213-216
    assert R.cosine_similarity([1, 0], [-1, 0], clamp=False) == pytest.approx(-1.0)
    title = R.cosine_similarity([0.8, 0.4, 0.2, 0.1], [0.75, 0.45, 0.25, 0.05])  # :31-47
    assert title == pytest.approx(0.9941180664180378, abs=1e-15)
    assert R.weighted_score(1.0, 1.0) == pytest.approx(1.0)                      # :124
    assert R.weighted_score(1.0, 0.0) == pytest.approx(0.7)                      # :187


def test_pure_python_restatement_matches_reference_outputs(golden):
    for e in golden["known_answers"] + golden["random_pairs"]:
        assert R.cosine_similarity(e["a"], e["b"]) == pytest.approx(e["clamped"], abs=1e-15)
        assert R.cosine_similarity(e["a"], e["b"], clamp=False) == pytest.approx(e["raw"], abs=1e-15)


def test_score_weighting_matches_reference(golden):
    for e in golden["score_weighting"]:
        title = R.cosine_similarity(e["vec_a"], e["vec_b"])
        svc = 1.0 if e["alert_service"] in e["incident_services"] else 0.0   # the golden cases are exact / disjoint
        assert R.weighted_score(title, svc) == pytest.approx(e["score"], abs=1e-12)


def test_numpy_restatement_matches_reference_outputs(golden):
    for e in golden["random_pairs"]:
        a = np.array(e["a"], dtype=np.float64)[None]
        b = np.array(e["b"], dtype=np.float64)[None]
        assert O.cosine_matrix(a, b)[0, 0] == pytest.approx(e["raw"], abs=1e-12)
        assert O.exact_cosine(a[0], b)[0] == pytest.approx(e["raw"], abs=1e-12)
        assert O.cosine_matrix(a, b, clamp=True)[0, 0] == pytest.approx(e["clamped"], abs=1e-12)


# ---- BASELINE.json config 1: 1 query x 1k docs, 384-d fp32, top-5 -------------------------
def _cfg1():
    C = np.random.default_rng(1001).standard_normal((1000, 384)).astype(np.float32)
    Q = np.random.default_rng(2001).standard_normal((1, 384)).astype(np.float32)
    return Q, C


def test_cfg1_topk_matches_reference_flat_scan(golden):
    g = golden["cfg1"]
    Q, C = _cfg1()
    ids, sc = O.cosine_topk(Q, C, g["k"])
    assert ids[0].tolist() == g["raw_ids"]
    assert np.allclose(sc[0], g["raw_scores"], atol=1e-7)
    ids_c, sc_c = O.cosine_topk(Q, C, g["k"], clamp=True)
    assert ids_c[0].tolist() == g["clamped_ids"]
    assert np.allclose(sc_c[0], g["clamped_scores"], atol=1e-7)
    # full-matrix checksum of all 1000 raw scores
    assert O.cosine_matrix(Q, C).sum() == pytest.approx(g["raw_all_scores_checksum"], abs=1e-9)
    # the pure-Python flat scan agrees as well
    pi, ps = R.topk_python([float(x) for x in Q[0]], [[float(x) for x in row] for row in C], g["k"], clamp=False)
    assert pi == g["raw_ids"]
    assert np.allclose(ps, g["raw_scores"], atol=1e-15)


# ---- semantics of the vectorised oracle -----------------------------------------------------
def test_tie_break_is_id_ascending_and_position_independent():
    Q, C = _cfg1()
    dup = np.concatenate([C, C[[813, 813, 526]]])            # bit-identical duplicates appended
    ids, sc = O.cosine_topk(Q, dup, 5)
    assert ids[0].tolist() == [813, 1000, 1001, 526, 1002]
    assert sc[0, 0] == sc[0, 1] == sc[0, 2]
    ext = np.arange(len(dup), dtype=np.int64)[::-1].copy()     # external ids reversed: ties follow ids, not rows
    ids2, _ = O.cosine_topk(Q, dup, 3, ids=ext)
    assert ids2[0].tolist() == sorted(ext[[813, 1000, 1001]].tolist())


def test_padding_zero_rows_and_tombstones():
    Q = np.array([[1.0, 0.0, 0.0]], dtype=np.float32)
    C = np.array([[1, 0, 0], [0, 0, 0], [-1, 0, 0], [0.5, 0.5, 0]], dtype=np.float32)
    ids, sc = O.cosine_topk(Q, C, 6)
    assert ids[0].tolist() == [0, 3, 1, 2, -1, -1]
    assert sc[0, 2] == 0.0                                     # zero norm -> 0.0 (similarity.py:94-95)
    assert np.isneginf(sc[0, 4:]).all()
    live = np.array([False, True, True, True])
    ids, _ = O.cosine_topk(Q, C, 2, live=live)
    assert ids[0].tolist() == [3, 1]
    ids, sc = O.cosine_topk(Q, C[:0], 3)
    assert (ids == -1).all() and np.isneginf(sc).all()


def test_tenant_scope_user_or_org():
    """weaviate_client.py:244-249: user_id == u OR org_id == o; org optional."""
    rng = np.random.default_rng(3)
    C = rng.standard_normal((200, 16)).astype(np.float32)
    Q = rng.standard_normal((3, 16)).astype(np.float32)
    ru = rng.integers(0, 4, 200).astype(np.int32)
    ro = rng.integers(-1, 3, 200).astype(np.int32)
    qu = np.array([0, 1, 2], dtype=np.int32)
    qo = np.array([-1, 2, 0], dtype=np.int32)
    ids, _ = O.cosine_topk(Q, C, 200, row_user=ru, row_org=ro, q_user=qu, q_org=qo)
    for i in range(3):
        got = set(int(x) for x in ids[i] if x >= 0)
        want = set(np.nonzero((ru == qu[i]) | ((qo[i] >= 0) & (ro == qo[i])))[0].tolist())
        assert got == want


def test_bf16_rounding_helper():
    x = np.array([1.0, 1.00390625, 1.005859375, -2.5, 3.14159265, 1e-30, 65504.0], dtype=np.float32)
    bits = O.f32_to_bf16_bits(x)
    back = O.bf16_bits_to_f32(bits)
    assert back[0] == 1.0 and back[3] == -2.5
    assert back[1] == 1.0            # 1 + 2^-8 is a tie: round to even (mantissa 0)
    assert back[2] == 1.0078125      # 1 + 1.5 * 2^-8 rounds up to 1 + 2^-7
    assert np.all(np.abs(back - x) <= np.abs(x) * 2.0 ** -8)
    assert np.array_equal(O.round_to_bf16(back), back)          # idempotent


def test_flat_search_port_agrees_with_oracle_on_separated_data():
    rng = np.random.default_rng(5)
    C = rng.standard_normal((3000, 64)).astype(np.float32)
    Q = rng.standard_normal((8, 64)).astype(np.float32)
    oi, osc = O.cosine_topk(Q, C, 10)
    fi, fsc = O.flat_search_f32(Q, C, 10)
    assert np.array_equal(oi, fi)
    assert np.allclose(osc, fsc, atol=1e-5)
