"""Row-sharded search over up to 8 GPUs: one process per GPU (torch.distributed), queries
replicated, corpus rows block-partitioned, one all-gather of per-shard top-k candidates and
a device-side merge.

The candidates that cross the wire are (fp64 score, int64 id) pairs: the shard-local results
are already exact (fp64 re-rank), so the merge is a pure (score desc, id asc) selection and the
sharded answer is identical to the single-GPU one.

torch is plumbing here (device tensors, streams, NCCL); the kernels are reached through the C
ABI.  `local_search` / `merge` are injectable so the protocol can be exercised on CPU with gloo
(tests/test_sharded_gloo.py).
"""

from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) owned by `rank`: equal blocks, the last rank takes the remainder."""
    per = n_rows // world
    lo = rank * per
    hi = n_rows if rank == world - 1 else lo + per
    return lo, hi


class ShardedSearcher:
    """search(queries) -> (ids [nq,k] int64, scores [nq,k] float32), identical on every rank.

    local_search(q, k) -> (scores64 [nq,k] float64, ids [nq,k] int64) for this rank's shard
    (padded with -inf / -1); merge(all_scores64 [G,nq,k], all_ids [G,nq,k], k) -> (ids, scores).
    """

    def __init__(self, local_search: Callable, merge: Callable, dist=None, world: int = 1):
        self.local_search, self.merge, self.dist, self.world = local_search, merge, dist, world

    def search(self, queries, k: int):
        import torch

        s64, ids = self.local_search(queries, k)
        if self.world == 1:
            return self.merge(s64.unsqueeze(0), ids.unsqueeze(0), k)
        nq = s64.shape[0]
        # ONE all-gather: fp64 keys and int64 ids travel as the two planes of an 8-byte-word buffer
        # (gathered along dim 0, the layout both backends accept)
        pack = torch.stack((s64.contiguous().view(torch.int64), ids.contiguous()), dim=0).view(2 * nq, k)
        allp = torch.empty((self.world * 2 * nq, k), dtype=torch.int64, device=ids.device)
        self.dist.all_gather_into_tensor(allp, pack)
        allp = allp.view(self.world, 2, nq, k)
        return self.merge(allp[:, 0].contiguous().view(torch.float64), allp[:, 1].contiguous(), k)


def make_gpu_searcher(index, dist=None, world: int = 1, device: int = 0, stream: Optional[int] = None) -> ShardedSearcher:
    """Wire an aurora_b200.engine.Index (this rank's shard) into a ShardedSearcher."""
    import torch

    from .engine import merge_topk_dev

    dev = torch.device("cuda", device)

    def _stream() -> int:
        return stream if stream is not None else (torch.cuda.current_stream().cuda_stream or 1)

    def local_search(q, k):
        nq = q.shape[0]
        sc = torch.empty(nq, k, device=dev, dtype=torch.float32)
        ids = torch.empty(nq, k, device=dev, dtype=torch.int64)
        s64 = torch.empty(nq, k, device=dev, dtype=torch.float64)
        index.search_dev(q.data_ptr(), nq, k, sc.data_ptr(), ids.data_ptr(), s64.data_ptr(), stream=_stream())
        return s64, ids

    def merge(all_s, all_i, k):
        g, nq = all_s.shape[0], all_s.shape[1]
        out_s = torch.empty(nq, k, device=dev, dtype=torch.float32)
        out_i = torch.empty(nq, k, device=dev, dtype=torch.int64)
        merge_topk_dev(device, all_s.data_ptr(), all_i.data_ptr(), g, nq, k, out_s.data_ptr(), out_i.data_ptr(),
                       stream=_stream())
        return out_i, out_s

    return ShardedSearcher(local_search, merge, dist=dist, world=world)
