"""Streaming exact cosine top-k and a threaded flat CPU index.  TEST / BENCH INFRASTRUCTURE ONLY.

Same answer as ``oracle.cosine_topk.cosine_topk`` (the pinned restatement of the reference's arithmetic,
``server/services/correlation/strategies/similarity.py:84-98`` applied to a flat scan, ordered
``(score desc, id asc)``), produced a corpus chunk at a time so a benchmark can check a full-size result
(1M .. 100M rows) without holding the corpus on the host:

  * candidate selection per chunk with a threaded fp32 product (torch, all host cores), keeping k + slack per
    query together with the candidates' vectors;
  * the final order from ``exact_cosine`` (extended-precision accumulation), exactly like ``cosine_topk``.

The fp32 selection can only differ from the fp64 one for scores closer than ~1e-6; with slack >= 32 the kept
set always contains the true top-k for the seeded corpora used here, and tests/test_oracle_streaming.py holds this
module to ``cosine_topk`` element for element.

``FlatIndexF32`` is the CPU baseline timed by ``bench.py --impl reference``: what a CPU flat cosine index does --
vectors L2-normalised once at import (Weaviate normalises at import for the cosine metric), then per batch one
threaded sgemm and a threaded partial selection per chunk.
"""

from __future__ import annotations

import numpy as np

from .cosine_topk import PAD_ID, PAD_SCORE, exact_cosine


class StreamingTopk:
    def __init__(self, Q: np.ndarray, k: int, slack: int = 32):
        import torch

        self.Q = np.ascontiguousarray(Q, dtype=np.float32)
        self.k, self.keep = int(k), int(k) + int(slack)
        qn = np.linalg.norm(self.Q.astype(np.float64), axis=1)
        self._Qn = torch.from_numpy((self.Q / np.where(qn > 0, qn, 1.0)[:, None]).astype(np.float32))
        nq, d = self.Q.shape
        self._s = torch.full((nq, 0), -np.inf, dtype=torch.float32)
        self._id = torch.zeros((nq, 0), dtype=torch.int64)
        self._vec = torch.zeros((nq, 0, d), dtype=torch.float32)
        self.rows = 0

    def add_chunk(self, C: np.ndarray, ids: np.ndarray) -> None:
        """C [n, D] float32 (bf16-rounded values for a bf16 store), ids [n] int64 (unique across chunks)."""
        import torch

        if C.shape[0] == 0:
            return
        Ct = torch.from_numpy(np.ascontiguousarray(C, dtype=np.float32))
        idt = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64))
        cn = torch.linalg.vector_norm(Ct, dim=1)
        inv = torch.where(cn > 0, 1.0 / cn, torch.zeros_like(cn))
        s = (self._Qn @ Ct.T) * inv[None, :]
        kk = min(self.keep, s.shape[1])
        vals, idx = torch.topk(s, kk, dim=1)
        new_id = idt[idx]
        new_vec = Ct[idx]                                   # [nq, kk, D]
        s_all = torch.cat([self._s, vals], dim=1)
        id_all = torch.cat([self._id, new_id], dim=1)
        vec_all = torch.cat([self._vec, new_vec], dim=1)
        kk2 = min(self.keep, s_all.shape[1])
        top, sel = torch.topk(s_all, kk2, dim=1)
        self._s = top
        self._id = torch.gather(id_all, 1, sel)
        self._vec = torch.gather(vec_all, 1, sel[:, :, None].expand(-1, -1, vec_all.shape[2]))
        self.rows += int(C.shape[0])

    def finish(self):
        """(ids [nq,k] int64, scores [nq,k] float32), best first, (score desc, id asc), padded like cosine_topk."""
        nq = self.Q.shape[0]
        out_ids = np.full((nq, self.k), PAD_ID, dtype=np.int64)
        out_sc = np.full((nq, self.k), PAD_SCORE, dtype=np.float32)
        ids = self._id.numpy()
        vec = self._vec.numpy()
        for i in range(nq):
            if ids.shape[1] == 0:
                continue
            ex = exact_cosine(self.Q[i], vec[i])
            order = np.lexsort((ids[i], -ex))[: self.k]
            out_ids[i, : order.size] = ids[i][order]
            out_sc[i, : order.size] = ex[order].astype(np.float32)
        return out_ids, out_sc


def cosine_topk_streaming(Q, C, k: int, ids=None, chunk: int = 65536, slack: int = 32):
    """Convenience wrapper with cosine_topk's signature (unfiltered case)."""
    C = np.asarray(C)
    ids = np.arange(C.shape[0], dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
    st = StreamingTopk(Q, k, slack)
    for lo in range(0, C.shape[0], chunk):
        st.add_chunk(C[lo:lo + chunk], ids[lo:lo + chunk])
    return st.finish()


class FlatIndexF32:
    """Threaded fp32 flat cosine index on the host cores (the CPU arm of bench.py)."""

    def __init__(self, dim: int, threads: int = 0):
        import torch

        if threads:
            torch.set_num_threads(threads)
        self.dim, self._blocks, self.rows = int(dim), [], 0
        self.threads = torch.get_num_threads()

    def add(self, C) -> None:
        """Append rows (numpy or torch float32 [n, dim]); they are L2-normalised here, once."""
        import torch

        Ct = C if isinstance(C, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(C, dtype=np.float32))
        cn = torch.linalg.vector_norm(Ct, dim=1, keepdim=True)
        self._blocks.append(torch.where(cn > 0, Ct / cn, torch.zeros_like(Ct)))
        self.rows += int(Ct.shape[0])

    def search(self, Q, k: int):
        import torch

        Qt = Q if isinstance(Q, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(Q, dtype=np.float32))
        qn = torch.linalg.vector_norm(Qt, dim=1, keepdim=True)
        Qn = torch.where(qn > 0, Qt / qn, torch.zeros_like(Qt))
        best_s = best_i = None
        base = 0
        for blk in self._blocks:
            s = Qn @ blk.T
            vals, idx = torch.topk(s, min(k, s.shape[1]), dim=1)
            idx = idx + base
            base += blk.shape[0]
            if best_s is None:
                best_s, best_i = vals, idx
            else:
                s2 = torch.cat([best_s, vals], dim=1)
                i2 = torch.cat([best_i, idx], dim=1)
                best_s, sel = torch.topk(s2, min(k, s2.shape[1]), dim=1)
                best_i = torch.gather(i2, 1, sel)
        return best_i.numpy(), best_s.numpy()
