"""Multi-GPU parity worker (launched by tests/test_gpu_multi.py under torchrun, one process per GPU): the
row-sharded search with BOTH cross-shard steps -- the fused peer-store exchange and the NCCL all-gather + merge --
must return, on every rank, exactly the oracle's top-k of the whole corpus (ids bit-exact, scores within 1e-3)."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from aurora_b200.engine import Index                     # noqa: E402
from aurora_b200.sharded import ShardedIndex, shard_bounds   # noqa: E402
from oracle import cosine_topk as O                      # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cases = [(50_000, 768, 256, 32, 6), (9_001, 1024, 300, 100, 3), (3_000, 384, 7, 5, 3)]
    for n, d, nq, k, iters in cases:
        rng = np.random.default_rng(n + d)
        C = O.round_to_bf16(rng.standard_normal((n, d)).astype(np.float32))
        C[n - 1] = C[1]                                   # an exact tie straddling the first and last shard
        lo, hi = shard_bounds(n, world, rank)
        with Index(d, max(hi - lo, 64), device=local) as ix:
            ix.add(C[lo:hi], np.arange(lo, hi, dtype=np.int64))
            for exchange in ("fused", "nccl"):
                sh = ShardedIndex(ix, dist, world, rank, local, nq_max=nq, k_max=k, exchange=exchange)
                for it in range(iters):                    # several rounds: both buffer parities, sequence numbers
                    Q = O.round_to_bf16(np.random.default_rng(1000 * it + n).standard_normal((nq, d)).astype(np.float32))
                    Q[0] = C[1]
                    q = torch.from_numpy(Q).to(dev).to(torch.bfloat16)
                    ids, sc = sh.search(q, k)
                    torch.cuda.synchronize()
                    oi, osc = O.cosine_topk(Q, C, k)
                    gi, gs = ids.cpu().numpy(), sc.cpu().numpy()
                    assert np.array_equal(gi, oi), f"rank {rank} {exchange} case {(n, d, nq, k)} it {it}: {int((gi != oi).sum())} id mismatches"
                    assert float(np.max(np.abs(gs - osc))) <= 1e-3
                    assert gi[0, 0] == 1 and gi[0, 1] == n - 1        # the tie resolves by id across shards
                if exchange == "fused":
                    done, status = sh._fx.status()
                    assert done == iters and status == 0, (done, status)
                sh.close()
        dist.barrier()
    if rank == 0:
        print("MGPU_OK", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
