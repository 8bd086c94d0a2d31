"""Pure-Python restatement of the only similarity arithmetic the reference owns.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``server/services/correlation/strategies/similarity.py:84-98``
(``SimilarityStrategy._cosine_similarity``):

* length mismatch or empty input            -> 0.0   (similarity.py:87-88)
* dot / (|a| * |b|) in Python floats (fp64)           (similarity.py:90-97)
* either norm zero                          -> 0.0   (similarity.py:94-95)
* result clamped into [0, 1]                          (similarity.py:98)

and the weighting of ``SimilarityStrategy.score`` (similarity.py:26-27,64):
``0.7 * title_sim + 0.3 * service_sim``.

Weaviate's cosine *distance* (1 - cos, range [0, 2]) is NOT clamped, and
``search_similar_good_rcas`` turns it back into ``similarity = 1 - distance``
without clamping (``server/routes/incident_feedback/weaviate_client.py:296-297``)
so ``clamp=False`` exposes the raw cosine as well.

Pinned by ``tests/test_oracle_golden.py`` against ``tests/golden/cosine_ref.json``
which ``oracle/gen_golden.py`` produced by importing the real reference function.
"""

from __future__ import annotations

import math
from typing import List, Sequence, Tuple

TITLE_WEIGHT = 0.7    # similarity.py:26
SERVICE_WEIGHT = 0.3  # similarity.py:27


def cosine_similarity(vec_a: Sequence[float], vec_b: Sequence[float], clamp: bool = True) -> float:
    n = len(vec_a)
    if n == 0 or n != len(vec_b):
        return 0.0
    # The reference reduces with the builtin sum(); keep the same reducer so the
    # interpreter's float-summation behaviour is shared, not re-implemented.
    ab = sum(vec_a[i] * vec_b[i] for i in range(n))
    aa = sum(x * x for x in vec_a)
    bb = sum(y * y for y in vec_b)
    na, nb = math.sqrt(aa), math.sqrt(bb)
    if na == 0 or nb == 0:
        return 0.0
    cos = ab / (na * nb)
    if not clamp:
        return cos
    return 0.0 if cos < 0.0 else (1.0 if cos > 1.0 else cos)


def topk_python(query: Sequence[float], corpus: Sequence[Sequence[float]], k: int,
                clamp: bool = True) -> Tuple[List[int], List[float]]:
    """Flat scan with the reference cosine, ordered by (score desc, row id asc).

    This is what BASELINE.json config 1 ("1 query x 1k-doc corpus, 384-d fp32,
    top-5 on CPU") looks like when only the reference's own arithmetic is used.
    O(N*D) Python: small cases only.
    """
    scored = [(-cosine_similarity(query, row, clamp=clamp), i) for i, row in enumerate(corpus)]
    scored.sort()
    top = scored[:k]
    return [i for _, i in top], [-s for s, _ in top]


def weighted_score(title_sim: float, service_sim: float) -> float:
    """similarity.py:64."""
    return TITLE_WEIGHT * title_sim + SERVICE_WEIGHT * service_sim
