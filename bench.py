#!/usr/bin/env python
"""Benchmark of the knowledge-base RAG hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1] -- batch of 256 queries against a
1M x 768 bf16 corpus, top-32, synthetic data (seeded randn), corpus resident in HBM.
A "step" = one batch of 256 queries through the fused similarity + top-k path.

  value     queries/s with the queries already in HBM (aur_search_dev on torch's stream),
            CUDA-event timed over K steps, max over ranks
  e2e       queries/s through the host-buffer C-ABI call (aur_search): pinned host queries
            -> H2D -> kernels -> D2H of scores + ids, every step
  roofline  dominant kernel (simtopk_tc): algorithmic bytes / its CUDA-event duration vs the
            measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the oracle's fp32 flat search ("port" of a CPU flat cosine index) on the host
            cores, bounded sample, rank 0, N=1 only

N > 1 (torchrun): the 1M-row corpus is row-sharded over the ranks (strong scaling); each
step = local search -> NCCL all-gather of (fp64 score, id) candidates -> device-side merge.

--impl reference: the reference's CPU path for the same config, timed on the host cores
(the Weaviate / t2v containers cannot run here; the oracle port restates the flat cosine
search, see oracle/cosine_topk.py).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS, DIM, NQ, TOPK = 1_000_000, 768, 256, 32
ALGO_BYTES = N_ROWS * DIM * 2 + NQ * DIM * 2 + NQ * TOPK * 8       # 1.5365e9 B / batch
METRIC = "RAG queries/sec (batch-256, 1M x 768 bf16, top-32)"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(r[2]) for r in self.rows if len(r) >= 7)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_flat_search_qps(sample_rows: int, reps: int = 1):
    """Oracle port: fp32 normalise + sgemm + select on all host cores; extrapolated linearly
    in the number of rows to the 1M-row batch."""
    from oracle import cosine_topk as O

    rng = np.random.default_rng(1002)
    C = rng.standard_normal((sample_rows, DIM), dtype=np.float32)
    Q = np.random.default_rng(2002).standard_normal((NQ, DIM), dtype=np.float32)
    O.flat_search_f32(Q[:8], C[:1000], TOPK)                      # warm BLAS threads
    t0 = time.perf_counter()
    for _ in range(reps):
        O.flat_search_f32(Q, C, TOPK)
    dt = (time.perf_counter() - t0) / reps
    return NQ / (dt * (N_ROWS / sample_rows)), dt


def cpu_pure_python_qps(rows: int = 1500):
    """The reference's own arithmetic (pure-Python cosine, similarity.py:84-98 restated in
    oracle/ref_cosine.py): one query against `rows` 768-d rows, extrapolated to 1M rows."""
    from oracle import ref_cosine as R

    rng = np.random.default_rng(7)
    C = [[float(x) for x in row] for row in rng.standard_normal((rows, DIM))]
    q = [float(x) for x in rng.standard_normal(DIM)]
    t0 = time.perf_counter()
    R.topk_python(q, C, TOPK, clamp=False)
    dt = time.perf_counter() - t0
    return 1.0 / (dt * (N_ROWS / rows))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = int(os.environ.get("AUR_BENCH_SAMPLE", "125000"))   # rows per step (the contract test shrinks it)
    for _ in range(args.warmup):
        cpu_flat_search_qps(sample)
    vals = [cpu_flat_search_qps(sample)[0] for _ in range(args.steps)]
    v = float(np.mean(vals))
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * NQ / v, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "batch-256 queries, 1M x 768 corpus, top-32 (BASELINE.json configs[1])",
                   "nq": NQ, "rows": N_ROWS, "dim": DIM, "k": TOPK},
        "cpu_baseline": {"value": v, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"256 queries x {sample} rows fp32 per step (numpy sgemm + argpartition), scaled x{N_ROWS // sample} to 1M rows",
                         "note": "the reference's Weaviate 1.27.6 / t2v containers cannot run here; this is the oracle's flat cosine search"},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch

    from aurora_b200 import _native as N
    from aurora_b200.engine import Index, merge_topk_packed_dev

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: aurora_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    # ---- corpus shard (rows [rank*n/G, (rank+1)*n/G)) generated on the device, queries replicated
    per = N_ROWS // world
    row_lo = rank * per
    n_local = per if rank < world - 1 else N_ROWS - row_lo
    ix = Index(DIM, n_local, dtype="bf16", device=local)
    # the C ABI reads NULL as "the index's own stream"; torch's default stream is the legacy
    # stream, whose explicit handle is cudaStreamLegacy (0x1)
    stream = torch.cuda.current_stream().cuda_stream or 1
    chunk = 125_000
    for lo in range(0, n_local, chunk):
        m = min(chunk, n_local - lo)
        g = torch.Generator(device=dev).manual_seed(1002 + (row_lo + lo) // chunk)
        rows = torch.randn(m, DIM, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
        ix.add_dev(rows.data_ptr(), m, np.arange(row_lo + lo, row_lo + lo + m, dtype=np.int64), stream=stream)
    gq = torch.Generator(device=dev).manual_seed(2002)
    q_dev = torch.randn(NQ, DIM, generator=gq, device=dev, dtype=torch.float32).to(torch.bfloat16)
    sc = torch.empty(NQ, TOPK, device=dev, dtype=torch.float32)
    # one 8-byte-word buffer per rank: plane 0 = fp64 ranking keys, plane 1 = int64 ids, so the
    # cross-shard exchange is a single all-gather
    pack = torch.empty(2, NQ, TOPK, device=dev, dtype=torch.int64)
    s64_ptr, ids_ptr = pack[0].data_ptr(), pack[1].data_ptr()
    ids = pack[1]
    if world > 1:
        gathered = torch.empty(world, 2, NQ, TOPK, device=dev, dtype=torch.int64)
        out_s = torch.empty(NQ, TOPK, device=dev, dtype=torch.float32)
        out_i = torch.empty(NQ, TOPK, device=dev, dtype=torch.int64)

    def step_dev():
        ix.search_dev(q_dev.data_ptr(), NQ, TOPK, sc.data_ptr(), ids_ptr, s64_ptr, stream=stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, pack)
            merge_topk_packed_dev(local, gathered.data_ptr(), world, NQ, TOPK, out_s.data_ptr(), out_i.data_ptr(), stream=stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()

    # ---- value: K steps, queries resident in HBM, CUDA events on the launching stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step_dev()
        e1.record()
        barrier()
        # keep the sampler running a little so short runs still get a few samples under load
        # (local search only: a time-bounded loop must not contain collectives)
        t_end = time.time() + 1.0
        while time.time() < t_end:
            ix.search_dev(q_dev.data_ptr(), NQ, TOPK, sc.data_ptr(), ids_ptr, s64_ptr, stream=stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = NQ / (ms * 1e-3)

    # ---- per-kernel time of the dominant kernel (library CUDA events on the same stream)
    kms, launches = [], 0
    for _ in range(min(args.steps, 20)):
        ix.search_dev(q_dev.data_ptr(), NQ, TOPK, sc.data_ptr(), ids_ptr, 0, stream=stream)
        torch.cuda.synchronize()
        st = ix.stats()
        kms.append(st["last_kernel_ms"])
        launches = st["last_launches"]
    kernel_ms = float(np.mean(kms))
    kernel_name = N.KERNEL_NAMES[st["last_kernel"]]

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region (every rank
    #      searches its shard; N > 1 adds the all-gather + merge and a D2H of the merged result)
    q_host = q_dev.cpu().pin_memory()
    h_sc = torch.empty(NQ, TOPK, dtype=torch.float32).pin_memory()
    h_id = torch.empty(NQ, TOPK, dtype=torch.int64).pin_memory()
    import ctypes as C

    lib = N.load()

    def step_e2e():
        if world == 1:
            N.check(lib.aur_search(ix._h, C.c_void_p(q_host.data_ptr()), NQ, TOPK, None, None,
                                   C.c_void_p(h_sc.data_ptr()), C.c_void_p(h_id.data_ptr())))
        else:
            q_dev.copy_(q_host, non_blocking=True)
            step_dev()
            h_sc.copy_(out_s, non_blocking=True)
            h_id.copy_(out_i, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = NQ / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    shard_bytes = n_local * DIM * 2 + NQ * DIM * 2 + NQ * TOPK * 8
    achieved = shard_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            traffic = json.load(open(tpath))["dram_bytes_per_launch"]
        except Exception:
            traffic = None
    out = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "batch-256 queries, 1M x 768 bf16 corpus, top-32 (BASELINE.json configs[1])",
                   "nq": NQ, "rows": N_ROWS, "dim": DIM, "k": TOPK, "parallelism": f"row-shard x{world}",
                   "l2": "corpus (1.5 GB) is larger than L2 (126 MB): no flush needed",
                   "kernel": kernel_name, "exchange": None if world == 1 else "one NCCL all-gather of the packed (fp64 score, id) planes + device merge"},
        "e2e": {"value": e2e, "unit": "queries/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": NQ * DIM * 2, "d2h_bytes_per_step": NQ * TOPK * 12},
        "gpu_launches": launches * args.steps + (args.steps if world > 1 else 0),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "simtopk_tc_kernel", "kernel_ms": kernel_ms,
                     "algorithmic_bytes": shard_bytes, "peak_source": peak_src},
        "clocks": clk.summary(),
    }
    if world == 1:
        cores = os.cpu_count() or 1
        sample = 250_000
        v, dt = cpu_flat_search_qps(sample)
        out["cpu_baseline"] = {
            "value": v, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"256 queries x {sample} rows fp32, one pass ({dt:.2f} s), scaled x{N_ROWS // sample} to 1M rows",
            "reference_pure_python_queries_per_s": cpu_pure_python_qps(),
            "note": "Weaviate 1.27.6 / t2v containers cannot run here; oracle fp32 flat cosine search on all host cores",
        }
        if not args.no_encoder:
            ix.close()
            torch.cuda.empty_cache()
            out["encoder"] = encoder_leg(local)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def encoder_leg(device: int) -> dict:
    """Second hot-path row (BASELINE.json configs[2] shape): bge-base-en
    dimensions, random-init bf16 weights, cfg3 chunk lengths ~N(384, 96) clipped to [16, 512].
    Device time of the forward (CUDA events inside the library), the same call end to end with
    host token ids in / host vectors out, and transformers' BertModel on the host cores beside it."""
    from aurora_b200.encoder import Encoder, EncoderConfig

    cfg = EncoderConfig()
    rng = np.random.default_rng(1003)
    n_seq = 192
    lens = np.clip(np.rint(rng.normal(384, 96, n_seq)), 16, 512).astype(np.int64)
    cu = np.zeros(n_seq + 1, np.int32)
    cu[1:] = np.cumsum(lens)
    tok = rng.integers(1000, cfg.vocab, size=int(cu[-1])).astype(np.int32)
    tok[cu[:-1]] = 101
    tok[cu[1:] - 1] = 102
    h, i = cfg.hidden, cfg.inter
    shapes = {"word_emb": (cfg.vocab, h), "pos_emb": (cfg.max_pos, h), "type_emb": (cfg.type_vocab, h),
              "emb_ln_g": (h,), "emb_ln_b": (h,)}
    for l in range(cfg.layers):
        for k, shp in {"wqkv": (3 * h, h), "bqkv": (3 * h,), "wo": (h, h), "bo": (h,), "ln1_g": (h,), "ln1_b": (h,),
                       "wi": (i, h), "bi": (i,), "wo2": (h, i), "bo2": (h,), "ln2_g": (h,), "ln2_b": (h,)}.items():
            shapes[f"l{l}.{k}"] = shp
    wrng = np.random.default_rng(7)
    with Encoder(cfg, max_tokens=int(cu[-1]) + 256, max_seqs=n_seq, device=device) as enc:
        for name, shp in shapes.items():
            a = (1.0 + 0.1 * wrng.standard_normal(shp)) if name.endswith("_g") else 0.02 * wrng.standard_normal(shp)
            enc.load_weights({name: a.astype(np.float32)})
        for _ in range(3):
            enc.encode_packed(tok, cu)
        dev_ms, e2e_ms = [], []
        for _ in range(10):
            t0 = time.perf_counter()
            enc.encode_packed(tok, cu)
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
            dev_ms.append(enc.stats()["total_ms"])
        st = enc.stats()
    ms, ems = float(np.median(dev_ms)), float(np.median(e2e_ms))
    flops = st["gemm_flops"] + st["attn_flops"]
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak_tf, peak_src = float(json.load(f)["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        peak_tf, peak_src = 1500.0, "fallback (B200_PROFILING.md)"
    out = {
        "workload": "bge-base-en dims (H768 L12 A12 I3072), random-init bf16, 192 chunks, lengths ~N(384,96) in [16,512]",
        "chunks": n_seq, "tokens": int(st["tokens"]), "ms_per_batch": ms, "chunks_per_s": n_seq / (ms * 1e-3),
        "tokens_per_s": st["tokens"] / (ms * 1e-3), "gpu_launches_per_batch": int(st["launches"]),
        "flops_per_batch": flops, "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak_tf,
                                               "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / peak_tf,
                                               "peak_source": peak_src,
                                               "note": "whole forward (GEMMs + attention + LayerNorm + pooling) over real (unpadded) tokens"},
        "e2e": {"chunks_per_s": n_seq / (ems * 1e-3), "ms_per_batch": ems, "h2d_bytes_per_batch": int(tok.nbytes + cu.nbytes),
                "d2h_bytes_per_batch": n_seq * h * 4},
    }
    try:   # host baseline: the class the reference's t2v sidecar runs, fp32, all host cores
        import torch
        from transformers import BertConfig as HFConfig, BertModel

        n_cpu = 16
        hf = BertModel(HFConfig(vocab_size=cfg.vocab, hidden_size=h, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                                intermediate_size=i, max_position_embeddings=cfg.max_pos), add_pooling_layer=False).eval()
        smax = int(lens[:n_cpu].max())
        ids = np.zeros((n_cpu, smax), np.int64)
        mask = np.zeros((n_cpu, smax), np.int64)
        for s_ in range(n_cpu):
            ids[s_, :lens[s_]] = tok[cu[s_]:cu[s_ + 1]]
            mask[s_, :lens[s_]] = 1
        with torch.no_grad():
            hf(input_ids=torch.from_numpy(ids[:2]), attention_mask=torch.from_numpy(mask[:2]))
            t0 = time.perf_counter()
            hf(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask))
            dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "chunks/s", "cores": os.cpu_count() or 1, "kind": "reference-stack",
                               "sample": f"{n_cpu} chunks, transformers.BertModel fp32 (torch {torch.get_num_threads()} threads), one padded batch ({dt:.2f} s)"}
    except Exception as e:   # transformers missing: report, do not fail the search line
        out["cpu_baseline"] = {"unavailable": str(e)[:120]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-encoder", action="store_true", help="skip the encoder leg of the N=1 run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
