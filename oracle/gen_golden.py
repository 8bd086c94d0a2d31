#!/usr/bin/env python
"""Generate tests/golden/*.json by running the REAL reference code.

Run in the authoring container only (``/root/reference`` does not exist on the GPU
box):  ``python oracle/gen_golden.py``.  The outputs are committed; nothing at test
time reads ``/root/reference``.

What is imported from the reference (read-only, no bytecode written):
  server/services/correlation/strategies/similarity.py  -> SimilarityStrategy
    ._cosine_similarity (:84-98), ._service_similarity (:124-153), .score (:29-64)

Golden content
  cosine_ref.json
    known_answers : the vectors the reference's own tests pin
                    (server/tests/services/correlation/test_similarity_strategy.py:193-216, :31-47)
    random_pairs  : seeded random pairs (dims 4..1024, some negative / zero / tiny)
                    with the reference's clamped result and the raw (unclamped) cosine
                    computed with the reference's formula
    cfg1          : BASELINE.json config 1 -- 1 query x 1k docs, 384-d fp32, top-5 --
                    flat scan with the reference function + sorted((-score, id))
    score_weighting: SimilarityStrategy.score() with a mocked embedding client
"""

from __future__ import annotations

import json
import math
import os
import sys
from unittest.mock import MagicMock, patch

import numpy as np

sys.dont_write_bytecode = True
REF = "/root/reference/server"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _raw_cosine_reference_formula(a, b):
    """Same expression as similarity.py:90-97 without the final clamp (:98)."""
    if len(a) != len(b) or len(a) == 0:
        return 0.0
    dot = sum(x * y for x, y in zip(a, b))
    na = math.sqrt(sum(x * x for x in a))
    nb = math.sqrt(sum(y * y for y in b))
    if na == 0 or nb == 0:
        return 0.0
    return dot / (na * nb)


def main() -> None:
    sys.path.insert(0, REF)
    from services.correlation.strategies.similarity import SimilarityStrategy  # noqa: E402

    cos = SimilarityStrategy._cosine_similarity
    os.makedirs(OUT, exist_ok=True)

    known = []
    for a, b in [([1, 2, 3], [1, 2, 3]), ([1, 0], [0, 1]), ([1, 0], [-1, 0]), ([], []),
                 ([1, 2], [1, 2, 3]), ([0.8, 0.4, 0.2, 0.1], [0.75, 0.45, 0.25, 0.05]),
                 ([0.5, 0.5, 0.5, 0.5], [0.5, 0.5, 0.5, 0.5]), ([0.9, 0.1, 0.0, 0.0], [0.0, 0.0, 0.1, 0.9]),
                 ([0.0, 0.0], [1.0, 2.0])]:
        known.append({"a": a, "b": b, "clamped": cos(a, b), "raw": _raw_cosine_reference_formula(a, b)})

    rng = np.random.default_rng(20260921)
    pairs = []
    for d in [4, 7, 16, 64, 100, 384, 768, 1024]:
        for variant in range(6):
            a = rng.standard_normal(d).astype(np.float32)
            b = rng.standard_normal(d).astype(np.float32)
            if variant == 1:
                b = (a + 0.05 * rng.standard_normal(d)).astype(np.float32)      # near-duplicate
            elif variant == 2:
                b = (-a).astype(np.float32)                                       # opposite -> clamp
            elif variant == 3:
                a = (a * 1e-20).astype(np.float32)                                # tiny magnitudes
            elif variant == 4:
                b = np.zeros(d, dtype=np.float32)                                 # zero norm
            al, bl = [float(x) for x in a], [float(x) for x in b]
            pairs.append({"a": al, "b": bl, "clamped": cos(al, bl), "raw": _raw_cosine_reference_formula(al, bl)})

    # BASELINE.json config 1 (SURVEY.md section 8(d) seeds)
    N, D, K = 1000, 384, 5
    C = np.random.default_rng(1001).standard_normal((N, D)).astype(np.float32)
    Q = np.random.default_rng(2001).standard_normal((1, D)).astype(np.float32)
    q = [float(x) for x in Q[0]]
    scored_clamped = sorted((-cos(q, [float(x) for x in C[i]]), i) for i in range(N))
    scored_raw = sorted((-_raw_cosine_reference_formula(q, [float(x) for x in C[i]]), i) for i in range(N))
    cfg1 = {
        "N": N, "D": D, "k": K, "corpus_seed": 1001, "query_seed": 2001,
        "clamped_ids": [i for _, i in scored_clamped[:K]],
        "clamped_scores": [-s for s, _ in scored_clamped[:K]],
        "raw_ids": [i for _, i in scored_raw[:K]],
        "raw_scores": [-s for s, _ in scored_raw[:K]],
        "raw_bottom_ids": [i for _, i in scored_raw[-K:]],
        "raw_all_scores_checksum": float(sum(-s for s, _ in scored_raw)),
    }

    # SimilarityStrategy.score weighting with a mocked embedding client (the reference's own
    # test technique, test_similarity_strategy.py:16-65)
    weighting = []
    strat = SimilarityStrategy()
    cases = [
        ([0.8, 0.4, 0.2, 0.1], [0.75, 0.45, 0.25, 0.05], "api-server", ["api-server"]),
        ([0.5, 0.5, 0.5, 0.5], [0.5, 0.5, 0.5, 0.5], "payment-service", ["payment-service"]),
        ([0.9, 0.1, 0.0, 0.0], [0.0, 0.0, 0.1, 0.9], "storage-node", ["network-gateway"]),
        ([1.0, 0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0], "svc", []),
    ]
    for va, vb, svc, inc in cases:
        with patch("services.correlation.strategies.similarity.get_embedding_client") as g:
            client = MagicMock()
            client.embed.side_effect = [va, vb]
            g.return_value = client
            val = strat.score("alert title", svc, "incident title", inc)
        weighting.append({"vec_a": va, "vec_b": vb, "alert_service": svc, "incident_services": inc, "score": val})

    with open(os.path.join(OUT, "cosine_ref.json"), "w") as f:
        json.dump({"generator": "oracle/gen_golden.py", "reference_commit": "e907d997",
                   "python": sys.version.split()[0],
                   "known_answers": known, "random_pairs": pairs, "cfg1": cfg1,
                   "score_weighting": weighting}, f)
    print("wrote", os.path.join(OUT, "cosine_ref.json"),
          f"({len(known)} known, {len(pairs)} random pairs, cfg1 top-{K})")


if __name__ == "__main__":
    main()
