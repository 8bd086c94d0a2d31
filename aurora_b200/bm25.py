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
VECTORISE_FROM = 256      # posting lists at least this long are scored as numpy arrays (cached per term), shorter ones in a Python loop


def tokenize(text: str) -> List[str]:
    return _WORD.findall(text.lower())


class BM25Index:
    """Inverted index over one text field per document id."""

    def __init__(self):
        self._postings: Dict[str, Dict[int, int]] = defaultdict(dict)   # term -> {doc id: tf}
        self._doc_terms: Dict[int, Counter] = {}
        self._doc_len: Dict[int, int] = {}
        self._total_len = 0
        self._arrays: Dict[str, tuple] = {}      # term -> (doc ids ascending, tf, doc lengths) for long posting lists; dropped on mutation

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
            self._arrays.pop(t, None)

    def remove(self, doc_id: int) -> bool:
        terms = self._doc_terms.pop(doc_id, None)
        if terms is None:
            return False
        self._total_len -= self._doc_len.pop(doc_id)
        for t in terms:
            plist = self._postings.get(t)
            self._arrays.pop(t, None)
            if plist is not None:
                plist.pop(doc_id, None)
                if not plist:
                    del self._postings[t]
        return True

    def _term_arrays(self, term: str, plist: Dict[int, int]):
        a = self._arrays.get(term)
        if a is None:
            import numpy as np

            ids = np.fromiter(plist.keys(), dtype=np.int64, count=len(plist))
            order = np.argsort(ids, kind="stable")
            ids = ids[order]
            tf = np.fromiter(plist.values(), dtype=np.float64, count=len(plist))[order]
            dl = np.fromiter((self._doc_len[int(d)] for d in ids), dtype=np.float64, count=len(ids))
            a = self._arrays[term] = (ids, tf, dl)
        return a

    def search(self, query: str, limit: int, allow: Optional[Callable[[int], bool]] = None,
               allowed: Optional[set] = None, allowed_sorted=None) -> List[Tuple[int, float]]:
        """Top-``limit`` (doc id, BM25 score), best first; ties by ascending id.  The pre-filter (tenant scope,
        metadata filter) comes as ``allowed`` -- a set of visible doc ids, intersected with every posting list at C
        speed (``allowed_sorted``: the same ids as an ascending int64 array, when the caller keeps one) -- or as a
        predicate ``allow``: documents outside do not exist for this query."""
        n_docs = len(self._doc_len)
        if n_docs == 0 or limit <= 0:
            return []
        avgdl = self._total_len / n_docs if self._total_len else 1.0
        scores: Dict[int, float] = defaultdict(float)
        doc_len = self._doc_len
        long_ids, long_contrib = [], []
        allowed_arr = allowed_sorted        # the same ids as `allowed`, ascending int64 (callers that keep one per tenant)
        for term in sorted(set(tokenize(query))):          # fixed order: equal indexes give bit-equal scores
            plist = self._postings.get(term)
            if not plist:
                continue
            idf = math.log(1.0 + (n_docs - len(plist) + 0.5) / (len(plist) + 0.5))
            if len(plist) >= VECTORISE_FROM and (allowed is None or len(allowed) >= VECTORISE_FROM):
                # a common word: its posting list is scored as arrays (a Python loop over 10^5 postings costs tens of
                # milliseconds per query; Weaviate's BM25 is native code)
                import numpy as np

                ids, tf, dl = self._term_arrays(term, plist)
                if allowed is not None:
                    if allowed_arr is None:
                        allowed_arr = np.fromiter(allowed, dtype=np.int64, count=len(allowed))
                        allowed_arr.sort()
                    if len(allowed_arr) <= len(ids):         # both ascending: look the shorter one up in the longer one
                        pos = np.minimum(np.searchsorted(ids, allowed_arr), len(ids) - 1)
                        pos = pos[ids[pos] == allowed_arr]
                        ids, tf, dl = ids[pos], tf[pos], dl[pos]
                    else:
                        pos = np.minimum(np.searchsorted(allowed_arr, ids), len(allowed_arr) - 1)
                        keep = allowed_arr[pos] == ids
                        ids, tf, dl = ids[keep], tf[keep], dl[keep]
                long_ids.append(ids)
                long_contrib.append(idf * tf * (K1 + 1.0) / (tf + K1 * (1.0 - B + B * dl / avgdl)))
                continue
            docs = plist.keys() & allowed if allowed is not None else plist.keys()
            for doc in docs:
                tf = plist[doc]
                scores[doc] += idf * tf * (K1 + 1.0) / (tf + K1 * (1.0 - B + B * doc_len[doc] / avgdl))
        if long_ids:
            import numpy as np

            if scores:                                      # the short lists' sums join as one more block, first in line
                long_ids.insert(0, np.fromiter(scores.keys(), dtype=np.int64, count=len(scores)))
                long_contrib.insert(0, np.fromiter(scores.values(), dtype=np.float64, count=len(scores)))
            all_ids = np.concatenate(long_ids)
            uniq, inv = np.unique(all_ids, return_inverse=True)
            total = np.bincount(inv, weights=np.concatenate(long_contrib), minlength=len(uniq))
            if allow is not None:
                keep = np.fromiter((bool(allow(int(d))) for d in uniq), dtype=bool, count=len(uniq))
                uniq, total = uniq[keep], total[keep]
            if len(uniq) > 4 * limit:                       # narrow before the exact (score desc, id asc) sort
                kth = np.partition(total, len(total) - limit)[len(total) - limit]
                keep = total >= kth
                uniq, total = uniq[keep], total[keep]
            order = np.lexsort((uniq, -total))[:limit]
            return [(int(uniq[i]), float(total[i])) for i in order]
        if allow is not None:
            scores = {d: v for d, v in scores.items() if allow(d)}
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
