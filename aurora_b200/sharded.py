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


# ----------------------------------------------------------------------------- fused exchange (no NCCL on the data path)
class FusedExchange:
    """Peer-mapped exchange buffers for the row-sharded search (include/aurora_b200.h, aur_exchange_*): every
    rank's finalize kernel stores its exact top-k rows straight into every rank's buffer over NVLink and the
    merge kernel that follows waits on delivery flags -- local search -> peer stores -> merge, three kernels on one
    stream, no collective call.  torch.distributed only carries the 64-byte IPC handles once, at set-up."""

    def __init__(self, dist, world: int, rank: int, device: int, nq_max: int, k_max: int):
        import ctypes as C

        from . import _native as N

        self._lib, self._N, self._dist = N.load(), N, dist
        self.world, self.rank, self.device = int(world), int(rank), int(device)
        self._h = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        N.check(self._lib.aur_exchange_create(self.device, self.rank, self.world, int(nq_max), int(k_max), C.byref(self._h), handle))
        if self.world > 1:
            mine = bytes(handle)
            gathered = [None] * self.world
            dist.all_gather_object(gathered, mine)
            blob = (C.c_uint8 * (64 * self.world)).from_buffer_copy(b"".join(gathered))
            N.check(self._lib.aur_exchange_connect(self._h, blob))
            dist.barrier()                      # every rank has mapped every buffer before anyone stores into one

    def search(self, index, q_ptr: int, nq: int, k: int, scores_ptr: int, ids_ptr: int, stream: int = 0) -> None:
        """All ranks, in lock step: on completion of the stream's work scores / ids hold the global top-k."""
        import ctypes as C

        self._N.check(self._lib.aur_search_exchange_dev(index._h, self._h, C.c_void_p(q_ptr), int(nq), int(k),
                                                        C.c_void_p(scores_ptr), C.c_void_p(ids_ptr), C.c_void_p(stream)))

    def status(self):
        import ctypes as C

        done, st = C.c_int64(0), C.c_int32(0)
        self._N.check(self._lib.aur_exchange_status(self._h, C.byref(done), C.byref(st)))
        return int(done.value), int(st.value)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            if self.world > 1:
                self._dist.barrier()            # nobody is still reading a peer's buffer
            self._lib.aur_exchange_close(self._h)
            self._h = None


class ShardedIndex:
    """This rank's shard + the cross-shard step, as bench.py and the multi-GPU tests drive it.
    ``search(q, k)`` returns (ids, scores) device tensors holding the GLOBAL top-k on every rank.
    exchange = "fused" (peer stores, default) or "nccl" (one all-gather of the packed (fp64 score, id) planes +
    device merge -- the baseline the fused path is measured against)."""

    def __init__(self, index, dist, world: int, rank: int, device: int, nq_max: int, k_max: int, exchange: str = "fused"):
        import torch

        self.index, self.dist, self.world, self.rank, self.device = index, dist, int(world), int(rank), int(device)
        self.exchange = exchange if world > 1 else "none"
        self._dev = torch.device("cuda", device)
        self._fx = FusedExchange(dist, world, rank, device, nq_max, k_max) if self.exchange == "fused" else None
        self._buf = {}

    def _bufs(self, nq: int, k: int):
        import torch

        key = (nq, k)
        if key not in self._buf:
            b = {"s": torch.empty(nq, k, device=self._dev, dtype=torch.float32),
                 "i": torch.empty(nq, k, device=self._dev, dtype=torch.int64)}
            if self.exchange == "nccl":
                b["pack"] = torch.empty(2, nq, k, device=self._dev, dtype=torch.int64)
                b["all"] = torch.empty(self.world, 2, nq, k, device=self._dev, dtype=torch.int64)
                b["ls"] = torch.empty(nq, k, device=self._dev, dtype=torch.float32)
            self._buf[key] = b
        return self._buf[key]

    def search(self, q, k: int, stream: Optional[int] = None):
        import torch

        from .engine import merge_topk_packed_dev

        s = stream if stream is not None else (torch.cuda.current_stream().cuda_stream or 1)
        nq = q.shape[0]
        b = self._bufs(nq, k)
        if self.exchange == "fused":
            self._fx.search(self.index, q.data_ptr(), nq, k, b["s"].data_ptr(), b["i"].data_ptr(), stream=s)
        elif self.exchange == "nccl":
            pack = b["pack"]
            self.index.search_dev(q.data_ptr(), nq, k, b["ls"].data_ptr(), pack[1].data_ptr(), pack[0].data_ptr(), stream=s)
            self.dist.all_gather_into_tensor(b["all"], pack)
            merge_topk_packed_dev(self.device, b["all"].data_ptr(), self.world, nq, k, b["s"].data_ptr(), b["i"].data_ptr(), stream=s)
        else:
            self.index.search_dev(q.data_ptr(), nq, k, b["s"].data_ptr(), b["i"].data_ptr(), stream=s)
        return b["i"], b["s"]

    def capture(self, q, k: int):
        """The whole step (local search -> peer stores -> merge) as ONE CUDA graph: ``replay()`` costs a single launch
        and the three kernels run back to back with no host in between.  Possible because every per-launch counter of
        the step (exchange-table epoch, exchange sequence number, candidate counts) lives in device memory and is
        advanced by the kernels themselves.  Returns (replay, ids, scores); all ranks must replay in lock step."""
        import torch

        if self.exchange == "nccl":
            raise RuntimeError("capture() covers the fused exchange and the single-GPU step, not the NCCL variant")
        side = torch.cuda.Stream(device=self._dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):                       # sizes every scratch buffer of this stream's search context
                self.search(q, k, stream=side.cuda_stream)
        side.synchronize()
        if self.dist is not None and self.world > 1:
            self.dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
            ids, sc = self.search(q, k, stream=side.cuda_stream)
        return g.replay, ids, sc

    def close(self) -> None:
        if self._fx is not None:
            self._fx.close()
            self._fx = None
