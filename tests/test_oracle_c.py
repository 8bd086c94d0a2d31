"""The C restatement of the oracle (oracle/cosine_topk.c) against the reference golden
vectors and the numpy oracle.  CPU only."""

import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import cosine_topk as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def orc():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle_cosine.so"))
    lib.orc_cosine.restype = C.c_double
    lib.orc_cosine.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.orc_topk.restype = None
    lib.orc_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def test_c_cosine_matches_reference_golden(orc):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "cosine_ref.json")))
    for e in g["known_answers"] + g["random_pairs"]:
        if len(e["a"]) != len(e["b"]):
            continue                                   # length mismatch is handled by the caller (similarity.py:87)
        a = np.array(e["a"], dtype=np.float64)
        b = np.array(e["b"], dtype=np.float64)
        p = lambda x: x.ctypes.data_as(C.c_void_p)     # noqa: E731
        assert orc.orc_cosine(p(a), p(b), len(a), 1) == pytest.approx(e["clamped"], abs=1e-12)
        assert orc.orc_cosine(p(a), p(b), len(a), 0) == pytest.approx(e["raw"], abs=1e-12)


def test_c_topk_matches_numpy_oracle_and_cfg1_golden(orc):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "cosine_ref.json")))["cfg1"]
    Cm = np.random.default_rng(1001).standard_normal((1000, 384)).astype(np.float32)
    Q = np.random.default_rng(2001).standard_normal((1, 384)).astype(np.float32)
    ids = np.empty((1, 5), np.int64)
    sc = np.empty((1, 5), np.float32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)         # noqa: E731
    orc.orc_topk(p(Q), p(Cm), None, 1, 1000, 384, 5, p(ids), p(sc))
    assert ids[0].tolist() == g["raw_ids"]
    assert np.allclose(sc[0], g["raw_scores"], atol=1e-6)
    rng = np.random.default_rng(4)
    Cm = rng.standard_normal((500, 32)).astype(np.float32)
    Cm[10] = Cm[400]
    Q = rng.standard_normal((6, 32)).astype(np.float32)
    ext = (np.arange(500, dtype=np.int64) * 7) % 501
    ids = np.empty((6, 12), np.int64)
    sc = np.empty((6, 12), np.float32)
    orc.orc_topk(p(Q), p(Cm), p(ext), 6, 500, 32, 12, p(ids), p(sc))
    oi, os_ = O.cosine_topk(Q, Cm, 12, ids=ext)
    assert np.array_equal(ids, oi) and np.allclose(sc, os_, atol=1e-6)
