"""World-size-2 (gloo, CPU) test of the row-sharded search protocol (aurora_b200/sharded.py):
shard bounds, the all-gather of (fp64 score, id) candidates and the merge rule.  The shard-local
search and the merge are CPU stand-ins built on the oracle (tests may use it); the GPU versions
of the same two callables are covered by tests/test_gpu_search.py::test_two_shard_merge."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aurora_b200.sharded import ShardedSearcher, shard_bounds
from oracle import cosine_topk as O

N_ROWS, DIM, NQ, K = 1037, 64, 9, 8          # odd row count: the last rank takes the remainder


def _corpus():
    rng = np.random.default_rng(1234)
    c = O.round_to_bf16(rng.standard_normal((N_ROWS, DIM)).astype(np.float32))
    q = O.round_to_bf16(rng.standard_normal((NQ, DIM)).astype(np.float32))
    c[700] = c[3]                              # an exact tie straddling the two shards
    c[5] = 0.0                                 # zero-norm row: cosine 0.0 (similarity.py:93-95)
    return q, c


def _local_search_factory(c_shard, row_lo):
    def local_search(q, k):
        qn = q.numpy()
        s = O.cosine_matrix(qn, c_shard)       # fp64
        ids = np.full((qn.shape[0], k), -1, np.int64)
        sc = np.full((qn.shape[0], k), -np.inf, np.float64)
        for i in range(qn.shape[0]):
            order = np.lexsort((np.arange(s.shape[1]), -s[i]))[:k]
            ids[i, :len(order)] = order + row_lo
            sc[i, :len(order)] = s[i, order]
        return torch.from_numpy(sc), torch.from_numpy(ids)
    return local_search


def _merge(all_s, all_i, k):
    g, nq, kk = all_s.shape
    s = all_s.permute(1, 0, 2).reshape(nq, g * kk).numpy()
    ids = all_i.permute(1, 0, 2).reshape(nq, g * kk).numpy()
    out_i = np.full((nq, k), -1, np.int64)
    out_s = np.full((nq, k), -np.inf, np.float32)
    for i in range(nq):
        valid = np.nonzero(ids[i] >= 0)[0]
        order = valid[np.lexsort((ids[i, valid], -s[i, valid]))][:k]
        out_i[i, :len(order)] = ids[i, order]
        out_s[i, :len(order)] = s[i, order].astype(np.float32)
    return torch.from_numpy(out_i), torch.from_numpy(out_s)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q, c = _corpus()
        lo, hi = shard_bounds(N_ROWS, world, rank)
        searcher = ShardedSearcher(_local_search_factory(c[lo:hi], lo), _merge, dist=dist, world=world)
        ids, sc = searcher.search(torch.from_numpy(q), K)
        np.save(os.path.join(out_dir, f"ids{rank}.npy"), ids.numpy())
        np.save(os.path.join(out_dir, f"sc{rank}.npy"), sc.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_cover_all_rows():
    for n in (0, 1, 7, 1037, 1_000_000):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])     # 3: shards of unequal size
def test_multi_rank_gloo_matches_single_process(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    q, c = _corpus()
    want_ids, want_sc = O.cosine_topk(q, c, K)
    for r in range(world):                     # every rank holds the same merged answer
        ids = np.load(tmp_path / f"ids{r}.npy")
        sc = np.load(tmp_path / f"sc{r}.npy")
        np.testing.assert_array_equal(ids, want_ids)
        np.testing.assert_allclose(sc, want_sc, rtol=0, atol=1e-6)


def test_single_rank_is_passthrough():
    q, c = _corpus()
    s = ShardedSearcher(_local_search_factory(c, 0), _merge, world=1)
    ids, sc = s.search(torch.from_numpy(q), K)
    want_ids, _ = O.cosine_topk(q, c, K)
    np.testing.assert_array_equal(ids.numpy(), want_ids)
