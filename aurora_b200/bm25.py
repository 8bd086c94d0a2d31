"""Sparse leg of the hybrid query + ranked fusion.

The reference asks Weaviate for ``collection.query.hybrid(query, alpha, fusion_type=HybridFusion.RANKED)``
(server/routes/knowledge_base/weaviate_client.py:252-259): a BM25F keyword search and a vector search
whose ranked lists are fused as ``sum_legs weight_leg / (rank + 60)`` with ``weight = alpha`` for the
vector leg and ``1 - alpha`` for the keyword leg.  Both live inside the Weaviate server (Go, CPU) --
not in /root/reference -- so this restates Weaviate 1.27's documented behaviour: BM25 with k1 = 1.2,
b = 0.75, "word" tokenisation (lower-case, split on non-alphanumerics), idf = ln(1 + (N - n + 0.5) /
(n + 0.5)); rankedFusion with the constant 60 and 0-based ranks.  Unpinned: the reference has no test
at this boundary.  Like in the reference this leg is host (CPU) work over text;
only the dense leg touches vectors and it runs on the GPU.
"""

from __future__ import annotations

import math
import re
from collections import Counter, defaultdict
from typing import Callable, Dict, Iterable, List, Optional, Tuple

_WORD = re.compile(r"[0-9a-z]+")
K1, B, RANK_CONSTANT = 1.2, 0.75, 60.0


def tokenize(text: str) -> List[str]:
    return _WORD.findall(text.lower())


class BM25Index:
    """Inverted index over one text field per document id."""

    def __init__(self):
        self._postings: Dict[str, Dict[int, int]] = defaultdict(dict)   # term -> {doc id: tf}
        self._doc_terms: Dict[int, Counter] = {}
        self._doc_len: Dict[int, int] = {}
        self._total_len = 0

    def __len__(self) -> int:
        return len(self._doc_len)

    def add(self, doc_id: int, text: str) -> None:
        if doc_id in self._doc_len:
            self.remove(doc_id)                       # upsert
        terms = Counter(tokenize(text))
        self._doc_terms[doc_id] = terms
        n = sum(terms.values())
        self._doc_len[doc_id] = n
        self._total_len += n
        for t, tf in terms.items():
            self._postings[t][doc_id] = tf

    def remove(self, doc_id: int) -> bool:
        terms = self._doc_terms.pop(doc_id, None)
        if terms is None:
            return False
        self._total_len -= self._doc_len.pop(doc_id)
        for t in terms:
            plist = self._postings.get(t)
            if plist is not None:
                plist.pop(doc_id, None)
                if not plist:
                    del self._postings[t]
        return True

    def search(self, query: str, limit: int, allow: Optional[Callable[[int], bool]] = None,
               allowed: Optional[set] = None) -> List[Tuple[int, float]]:
        """Top-``limit`` (doc id, BM25 score), best first; ties by ascending id.  The pre-filter (tenant scope,
        metadata filter) comes as ``allowed`` -- a set of visible doc ids, intersected with every posting list at C
        speed -- or as a predicate ``allow``: documents outside do not exist for this query."""
        n_docs = len(self._doc_len)
        if n_docs == 0 or limit <= 0:
            return []
        avgdl = self._total_len / n_docs if self._total_len else 1.0
        scores: Dict[int, float] = defaultdict(float)
        doc_len = self._doc_len
        for term in set(tokenize(query)):
            plist = self._postings.get(term)
            if not plist:
                continue
            idf = math.log(1.0 + (n_docs - len(plist) + 0.5) / (len(plist) + 0.5))
            docs = plist.keys() & allowed if allowed is not None else plist.keys()
            for doc in docs:
                if allow is not None and not allow(doc):
                    continue
                tf = plist[doc]
                scores[doc] += idf * tf * (K1 + 1.0) / (tf + K1 * (1.0 - B + B * doc_len[doc] / avgdl))
        return sorted(scores.items(), key=lambda kv: (-kv[1], kv[0]))[:limit]


def ranked_fusion(legs: Iterable[Tuple[float, List[int]]], limit: int) -> List[Tuple[int, float]]:
    """Weaviate rankedFusion: ``legs`` = (weight, ids best-first); score(id) = sum over legs of
    weight / (rank + 60), rank 0-based.  Returns the top ``limit`` (id, fused score), ties by id."""
    fused: Dict[int, float] = defaultdict(float)
    for weight, ids in legs:
        if weight <= 0.0:
            continue
        for rank, doc in enumerate(ids):
            fused[doc] += weight / (rank + RANK_CONSTANT)
    return sorted(fused.items(), key=lambda kv: (-kv[1], kv[0]))[:limit]
