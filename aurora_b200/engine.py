"""Python handle over one corpus shard in HBM (the C ABI of include/aurora_b200.h).

numpy in / numpy out for the host entry points; ``*_dev`` methods take raw device
pointers (``tensor.data_ptr()``) so torch is only plumbing for memory and streams.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _native as N


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit patterns (uint16), round-to-nearest-even.  Host-side format
    conversion of inputs only; no scoring happens on the CPU."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
    return (r >> np.uint32(16)).astype(np.uint16)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Index:
    """One row-shard of the corpus, resident on one GPU."""

    def __init__(self, dim: int, capacity: int, dtype: str = "bf16", device: int = 0):
        self._lib = N.load()
        self.dim, self.capacity, self.device = int(dim), int(capacity), int(device)
        self.dtype = {"bf16": N.AUR_BF16, "f32": N.AUR_F32}[dtype]
        self._h = C.c_void_p()
        cfg = N.AurConfig(device=self.device, dim=self.dim, dtype=self.dtype, reserved=0, capacity=self.capacity)
        N.check(self._lib.aur_open(C.byref(cfg), C.byref(self._h)))

    # -------------------------------------------------------------- lifecycle
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.aur_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -------------------------------------------------------------- helpers
    def _rows_buffer(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x)
        if x.ndim != 2 or x.shape[1] != self.dim:
            raise ValueError(f"expected [n, {self.dim}] rows, got {x.shape}")
        if self.dtype == N.AUR_BF16:
            if x.dtype == np.uint16:
                return np.ascontiguousarray(x)
            return to_bf16_bits(x)
        return np.ascontiguousarray(x, dtype=np.float32)

    # -------------------------------------------------------------- ingest
    def add(self, rows: np.ndarray, ids: np.ndarray, user_codes: Optional[np.ndarray] = None,
            org_codes: Optional[np.ndarray] = None) -> None:
        buf = self._rows_buffer(rows)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        if ids.shape != (buf.shape[0],):
            raise ValueError("ids must be [n]")
        u = None if user_codes is None else np.ascontiguousarray(user_codes, dtype=np.int32)
        o = None if org_codes is None else np.ascontiguousarray(org_codes, dtype=np.int32)
        N.check(self._lib.aur_add(self._h, _ptr(buf), _ptr(ids), _ptr(u), _ptr(o), buf.shape[0]))

    def add_dev(self, rows_ptr: int, n: int, ids: np.ndarray, user_codes=None, org_codes=None, stream: int = 0) -> None:
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        u = None if user_codes is None else np.ascontiguousarray(user_codes, dtype=np.int32)
        o = None if org_codes is None else np.ascontiguousarray(org_codes, dtype=np.int32)
        N.check(self._lib.aur_add_dev(self._h, C.c_void_p(rows_ptr), _ptr(ids), _ptr(u), _ptr(o), int(n),
                                      C.c_void_p(stream)))

    def remove(self, ids) -> int:
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        removed = C.c_int64(0)
        N.check(self._lib.aur_remove(self._h, _ptr(ids), ids.shape[0], C.byref(removed)))
        return int(removed.value)

    # -------------------------------------------------------------- snapshot
    def export(self):
        """All appended rows in append order: (rows [n, dim] uint16 bf16 bits or float32, ids, user codes,
        org codes, live mask).  Tombstoned rows are included with live = False."""
        n = self.stats()["rows"]
        rows = np.empty((n, self.dim), dtype=np.uint16 if self.dtype == N.AUR_BF16 else np.float32)
        ids = np.empty(n, dtype=np.int64)
        user, org = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        live = np.empty(n, dtype=np.uint8)
        N.check(self._lib.aur_export(self._h, _ptr(rows), _ptr(ids), _ptr(user), _ptr(org), _ptr(live), n))
        return rows, ids, user, org, live.astype(bool)

    def save(self, path: str) -> None:
        """Snapshot of the live rows (tombstones are compacted away) as one .npz file."""
        rows, ids, user, org, live = self.export()
        np.savez(path, rows=rows[live], ids=ids[live], user=user[live], org=org[live], dim=self.dim,
                 dtype=self.dtype, capacity=self.capacity)

    @classmethod
    def load(cls, path: str, capacity: Optional[int] = None, device: int = 0) -> "Index":
        z = np.load(path if path.endswith(".npz") else path + ".npz")
        dtype = "bf16" if int(z["dtype"]) == N.AUR_BF16 else "f32"
        ix = cls(int(z["dim"]), int(capacity or z["capacity"]), dtype=dtype, device=device)
        if len(z["ids"]):
            ix.add(z["rows"], z["ids"], z["user"], z["org"])
        return ix

    # -------------------------------------------------------------- search
    def search(self, queries: np.ndarray, k: int, q_user: Optional[np.ndarray] = None,
               q_org: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
        """Host buffers in, host buffers out: (ids [nq,k] int64, scores [nq,k] float32)."""
        q = self._rows_buffer(queries)
        nq = q.shape[0]
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        u = None if q_user is None else np.ascontiguousarray(q_user, dtype=np.int32)
        o = None if q_org is None else np.ascontiguousarray(q_org, dtype=np.int32)
        N.check(self._lib.aur_search(self._h, _ptr(q), nq, int(k), _ptr(u), _ptr(o), _ptr(scores), _ptr(ids)))
        return ids, scores

    def search_snapshot(self, queries: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray, int]:
        """(ids, scores, snapshot_rows): the search together with the number of appended rows it saw -- with a
        concurrent writer the answer is the top-k of exactly that prefix."""
        q = self._rows_buffer(queries)
        nq = q.shape[0]
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        snap = C.c_int64(-1)
        N.check(self._lib.aur_search_ex(self._h, _ptr(q), nq, int(k), None, None, _ptr(scores), _ptr(ids), C.byref(snap)))
        return ids, scores, int(snap.value)

    def read_rows(self, row0: int, n: int):
        """(rows [n, dim] uint16 bf16 bits or float32, ids [n]) of appended rows row0 .. row0 + n."""
        rows = np.empty((n, self.dim), dtype=np.uint16 if self.dtype == N.AUR_BF16 else np.float32)
        ids = np.empty(n, dtype=np.int64)
        N.check(self._lib.aur_read_rows(self._h, int(row0), int(n), _ptr(rows), _ptr(ids)))
        return rows, ids

    def compact(self) -> int:
        """Reclaim tombstoned rows (exclusive; waits for searches in flight).  Returns the rows freed."""
        freed = C.c_int64(0)
        N.check(self._lib.aur_compact(self._h, C.byref(freed)))
        return int(freed.value)

    def search_subset(self, queries: np.ndarray, k: int, allow_ids: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Search restricted to the rows whose ids are listed (a resolved metadata pre-filter): the other rows'
        inverse norms are masked on the device and the same kernels run."""
        q = self._rows_buffer(queries)
        nq = q.shape[0]
        allow = np.ascontiguousarray(allow_ids, dtype=np.int64)
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        N.check(self._lib.aur_search_subset(self._h, _ptr(q), nq, int(k), _ptr(allow), allow.shape[0], _ptr(scores), _ptr(ids)))
        return ids, scores

    def search_dev(self, q_ptr: int, nq: int, k: int, scores_ptr: int, ids_ptr: int, scores64_ptr: int = 0,
                   q_user_ptr: int = 0, q_org_ptr: int = 0, stream: int = 0) -> None:
        """Everything in HBM; asynchronous on ``stream`` (0 = the index's own stream)."""
        N.check(self._lib.aur_search_dev(self._h, C.c_void_p(q_ptr), int(nq), int(k), C.c_void_p(q_user_ptr),
                                         C.c_void_p(q_org_ptr), C.c_void_p(scores_ptr), C.c_void_p(ids_ptr),
                                         C.c_void_p(scores64_ptr), C.c_void_p(stream)))

    def debug_tc_scores(self, q_ptr: int, nq: int, cta_group: int, out_ptr: int, stream: int = 0) -> int:
        n = C.c_int32(0)
        N.check(self._lib.aur_debug_tc_scores(self._h, C.c_void_p(q_ptr), int(nq), int(cta_group),
                                              C.c_void_p(out_ptr), C.byref(n), C.c_void_p(stream)))
        return int(n.value)

    # -------------------------------------------------------------- misc
    def set_kernel(self, kernel: int) -> None:
        N.check(self._lib.aur_set_option(self._h, b"kernel", int(kernel)))

    def sync(self) -> None:
        N.check(self._lib.aur_sync(self._h))

    def stats(self) -> dict:
        st = N.AurStats()
        N.check(self._lib.aur_get_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in N.AurStats._fields_}


class MultiIndex:
    """One process, one shard per GPU of the host (the daemon's deployment on an 8-GPU box): the corpus is row-sharded
    by ``id mod n`` -- an upsert or a delete always lands on the shard that holds the old row -- every search runs on
    all shards at once (one host thread per device; the C calls release the GIL) and the per-shard top-k lists are
    merged on the host by (score desc, id asc), the order the device merge uses (csrc/kernels_simt.cu merge_topk_kernel).
    Same surface as ``Index`` for what ``retriever.KnowledgeBase`` and the daemon call.  The rank-per-GPU deployment with
    the fused peer-store exchange (sharded.ShardedIndex) stays the throughput path; this one trades ~0.1 ms of host
    merge for a single owner process.  ``shard_factory(dim, capacity, device)`` builds one shard (tests: a CPU double)."""

    def __init__(self, dim: int, capacity: int, devices=None, dtype: str = "bf16", shard_factory=None, _shards=None):
        from concurrent.futures import ThreadPoolExecutor

        if devices is None:
            devices = list(range(N.load().aur_device_count()))
        if not devices:
            raise RuntimeError("MultiIndex needs at least one device")
        self.dim, self.capacity, self.devices = int(dim), int(capacity), [int(d) for d in devices]
        n = len(self.devices)
        per = (self.capacity + n - 1) // n
        per += max(64, per // 8)                    # id mod n is balanced only statistically
        self._native_merge = shard_factory is None and _shards is None
        if shard_factory is None:
            shard_factory = lambda dim_, cap_, dev_: Index(dim_, cap_, dtype=dtype, device=dev_)   # noqa: E731
        self.shards = _shards if _shards is not None else [shard_factory(self.dim, per, d) for d in self.devices]
        self.dtype = getattr(self.shards[0], "dtype", N.AUR_BF16)
        self._pool = ThreadPoolExecutor(max_workers=n, thread_name_prefix="aurora-b200-shard")

    # -------------------------------------------------------------- plumbing
    def _each(self, fn):
        """fn(shard index, shard) on every shard concurrently; results in shard order."""
        if len(self.shards) == 1:
            return [fn(0, self.shards[0])]
        return list(self._pool.map(lambda t: fn(*t), enumerate(self.shards)))

    def _split(self, ids: np.ndarray):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        owner = np.mod(ids, len(self.shards))
        return ids, [np.nonzero(owner == s)[0] for s in range(len(self.shards))]

    def _merge(self, parts, k: int):
        if len(parts) == 1:
            return parts[0]
        if self._native_merge:      # k-way merge of the sorted lists in C (csrc/host_merge.cpp): microseconds
            ids = np.ascontiguousarray(np.stack([p[0] for p in parts]), dtype=np.int64)
            sc = np.ascontiguousarray(np.stack([p[1] for p in parts]), dtype=np.float32)
            nq, k_in = ids.shape[1], ids.shape[2]
            out_s, out_i = np.empty((nq, k), dtype=np.float32), np.empty((nq, k), dtype=np.int64)
            N.check(N.load().aur_merge_topk_host(_ptr(sc), _ptr(ids), len(parts), nq, k_in, k, _ptr(out_s), _ptr(out_i)))
            return out_i, out_s
        ids = np.concatenate([p[0] for p in parts], axis=1)         # CPU doubles in the tests: the same order in numpy
        sc = np.concatenate([p[1] for p in parts], axis=1)
        key_id = np.where(ids < 0, np.iinfo(np.int64).max, ids)      # empty slots last
        order = np.lexsort((key_id, -sc.astype(np.float64)), axis=1)[:, :k]
        return np.take_along_axis(ids, order, axis=1), np.take_along_axis(sc, order, axis=1)

    # -------------------------------------------------------------- ingest / deletes
    def add(self, rows: np.ndarray, ids: np.ndarray, user_codes=None, org_codes=None) -> None:
        rows = np.asarray(rows)
        ids, sel = self._split(ids)
        u = None if user_codes is None else np.asarray(user_codes, dtype=np.int32)
        o = None if org_codes is None else np.asarray(org_codes, dtype=np.int32)

        def put(s, shard):
            ix = sel[s]
            if len(ix):
                shard.add(rows[ix], ids[ix], None if u is None else u[ix], None if o is None else o[ix])
        self._each(put)

    def remove(self, ids) -> int:
        ids, sel = self._split(ids)
        return int(sum(self._each(lambda s, shard: shard.remove(ids[sel[s]]) if len(sel[s]) else 0)))

    # -------------------------------------------------------------- search
    def search(self, queries: np.ndarray, k: int, q_user=None, q_org=None):
        return self._merge(self._each(lambda s, shard: shard.search(queries, k, q_user, q_org)), k)

    def search_subset(self, queries: np.ndarray, k: int, allow_ids: np.ndarray):
        allow, sel = self._split(allow_ids)
        return self._merge(self._each(lambda s, shard: shard.search_subset(queries, k, allow[sel[s]])), k)

    # -------------------------------------------------------------- maintenance
    def stats(self) -> dict:
        per = self._each(lambda s, shard: shard.stats())
        out = {key: int(sum(p.get(key, 0) for p in per)) for key in ("rows", "live", "searches")}
        out["last_kernel"] = per[0].get("last_kernel", 0)
        out["shards"] = len(per)
        out["rows_per_shard"] = [int(p.get("rows", 0)) for p in per]
        return out

    def compact(self) -> int:
        return int(sum(self._each(lambda s, shard: shard.compact())))

    def sync(self) -> None:
        self._each(lambda s, shard: shard.sync())

    def save(self, path: str) -> None:
        """One .npz like Index.save (live rows of all shards): a snapshot restores onto any number of devices."""
        parts = self._each(lambda s, shard: shard.export())
        cat = lambda i: np.concatenate([np.asarray(p[i])[p[4]] for p in parts])   # noqa: E731
        np.savez(path, rows=cat(0), ids=cat(1), user=cat(2), org=cat(3), dim=self.dim, dtype=self.dtype, capacity=self.capacity)

    @classmethod
    def load(cls, path: str, capacity: Optional[int] = None, devices=None, shard_factory=None) -> "MultiIndex":
        z = np.load(path if path.endswith(".npz") else path + ".npz")
        dtype = "bf16" if int(z["dtype"]) == N.AUR_BF16 else "f32"
        mi = cls(int(z["dim"]), int(capacity or z["capacity"]), devices=devices, dtype=dtype, shard_factory=shard_factory)
        if len(z["ids"]):
            mi.add(z["rows"], z["ids"], z["user"], z["org"])
        return mi

    def close(self) -> None:
        for sh in self.shards:
            sh.close()
        self.shards = []
        self._pool.shutdown(wait=False)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def merge_topk_packed_dev(device: int, packed_ptr: int, n_shards: int, nq: int, k: int, out_scores_ptr: int,
                          out_ids_ptr: int, out_scores64_ptr: int = 0, stream: int = 0) -> None:
    """packed: [n_shards][2][nq][k] 8-byte words (plane 0 fp64 scores, plane 1 int64 ids)."""
    lib = N.load()
    N.check(lib.aur_merge_topk_packed_dev(int(device), C.c_void_p(packed_ptr), int(n_shards), int(nq), int(k),
                                          C.c_void_p(out_scores_ptr), C.c_void_p(out_ids_ptr),
                                          C.c_void_p(out_scores64_ptr), C.c_void_p(stream)))


def merge_topk_dev(device: int, in_scores64_ptr: int, in_ids_ptr: int, n_shards: int, nq: int, k: int,
                   out_scores_ptr: int, out_ids_ptr: int, out_scores64_ptr: int = 0, stream: int = 0) -> None:
    lib = N.load()
    N.check(lib.aur_merge_topk_dev(int(device), C.c_void_p(in_scores64_ptr), C.c_void_p(in_ids_ptr), int(n_shards),
                                   int(nq), int(k), C.c_void_p(out_scores_ptr), C.c_void_p(out_ids_ptr),
                                   C.c_void_p(out_scores64_ptr), C.c_void_p(stream)))


def cosine_pairs(a: np.ndarray, b: np.ndarray, clamp: bool = False, device: int = 0) -> np.ndarray:
    """Row-wise cosine on the GPU (fp64 accumulate).  Mirrors
    SimilarityStrategy._cosine_similarity (similarity.py:84-98) for batches."""
    lib = N.load()
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    if a.shape != b.shape or a.ndim != 2:
        raise ValueError("a and b must both be [n, dim]")
    out = np.empty(a.shape[0], dtype=np.float64)
    N.check(lib.aur_cosine_pairs(int(device), _ptr(a), _ptr(b), a.shape[0], a.shape[1], int(bool(clamp)), _ptr(out)))
    return out


class DeviceBuffer:
    """Raw HBM buffer owned through the C ABI (for callers without torch)."""

    def __init__(self, nbytes: int, device: int = 0):
        self._lib = N.load()
        self.device, self.nbytes = int(device), int(nbytes)
        p = C.c_void_p()
        N.check(self._lib.aur_dev_malloc(self.device, self.nbytes, C.byref(p)))
        self.ptr = int(p.value)

    def upload(self, a: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        if a.nbytes > self.nbytes:
            raise ValueError("buffer too small")
        N.check(self._lib.aur_memcpy_h2d(self.device, C.c_void_p(self.ptr), _ptr(a), a.nbytes))
        return self

    def download(self, a: np.ndarray) -> np.ndarray:
        if not a.flags["C_CONTIGUOUS"] or a.nbytes > self.nbytes:
            raise ValueError("need a contiguous array no larger than the buffer")
        N.check(self._lib.aur_memcpy_d2h(self.device, _ptr(a), C.c_void_p(self.ptr), a.nbytes))
        return a

    def free(self) -> None:
        if self.ptr:
            self._lib.aur_dev_free(self.device, C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
