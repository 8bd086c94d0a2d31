#!/usr/bin/env python
"""One encoder layer at bge-base dims for an ncu capture of its kernels (four GEMMs + attention):
  ncu --set full --clock-control none -k regex:"gemm_tc|attn_tc" --launch-skip 10 -c 5 -o out python tools/encoder_layer_prof.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from aurora_b200.encoder import Encoder, EncoderConfig

cfg = EncoderConfig(layers=1)
rng = np.random.default_rng(3)
n_seq = int(os.environ.get("AUR_NSEQ", "64"))
lens = np.clip(np.rint(rng.normal(384, 96, n_seq)), 16, 512).astype(np.int64)
cu = np.zeros(n_seq + 1, np.int32); cu[1:] = np.cumsum(lens)
tok = rng.integers(1000, cfg.vocab, int(cu[-1])).astype(np.int32)
enc = Encoder(cfg, max_tokens=int(cu[-1]) + 1024, max_seqs=n_seq, device=0)
enc.load_weights(bench.random_bert_weights(cfg, 7))
for _ in range(3):
    out = enc.encode_packed(tok, cu)
print("tokens", int(cu[-1]), "out", out.shape)
enc.close()
