"""Generate tests/golden/bert_ref.json from the installed transformers.BertModel (5.5.0): the
class the reference's t2v-transformers sidecar runs (embedding_client.py:52-59 posts to it).
Weights are the seeded random init of oracle/bert_encoder.init_weights loaded into the HF
module, inputs are right-padded with an attention mask, outputs are the pooled vectors.

    python -m oracle.gen_golden_bert          # run from the repo root, in the authoring container
"""

import json
import os

import numpy as np
import torch
from transformers import BertConfig as HFConfig, BertModel

from . import bert_encoder as B


def load_into_hf(cfg: B.BertConfig, w) -> BertModel:
    hf = HFConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers,
                  num_attention_heads=cfg.heads, intermediate_size=cfg.inter, max_position_embeddings=cfg.max_pos,
                  type_vocab_size=cfg.type_vocab, layer_norm_eps=cfg.ln_eps, hidden_act="gelu",
                  hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    m = BertModel(hf, add_pooling_layer=False).eval()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    sd = {
        "embeddings.word_embeddings.weight": t(w["word_emb"]),
        "embeddings.position_embeddings.weight": t(w["pos_emb"]),
        "embeddings.token_type_embeddings.weight": t(w["type_emb"]),
        "embeddings.LayerNorm.weight": t(w["emb_ln_g"]), "embeddings.LayerNorm.bias": t(w["emb_ln_b"]),
    }
    h = cfg.hidden
    for l in range(cfg.layers):
        p, q = f"l{l}.", f"encoder.layer.{l}."
        for i, nm in enumerate(("query", "key", "value")):
            sd[q + f"attention.self.{nm}.weight"] = t(w[p + "wqkv"][i * h:(i + 1) * h])
            sd[q + f"attention.self.{nm}.bias"] = t(w[p + "bqkv"][i * h:(i + 1) * h])
        sd[q + "attention.output.dense.weight"] = t(w[p + "wo"]); sd[q + "attention.output.dense.bias"] = t(w[p + "bo"])
        sd[q + "attention.output.LayerNorm.weight"] = t(w[p + "ln1_g"]); sd[q + "attention.output.LayerNorm.bias"] = t(w[p + "ln1_b"])
        sd[q + "intermediate.dense.weight"] = t(w[p + "wi"]); sd[q + "intermediate.dense.bias"] = t(w[p + "bi"])
        sd[q + "output.dense.weight"] = t(w[p + "wo2"]); sd[q + "output.dense.bias"] = t(w[p + "bo2"])
        sd[q + "output.LayerNorm.weight"] = t(w[p + "ln2_g"]); sd[q + "output.LayerNorm.bias"] = t(w[p + "ln2_b"])
    missing, unexpected = m.load_state_dict(sd, strict=False)
    missing = [k for k in missing if "position_ids" not in k and "token_type_ids" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return m


def hf_encode(cfg: B.BertConfig, w, tokens, cu) -> np.ndarray:
    m = load_into_hf(cfg, w).double()
    n_seq = len(cu) - 1
    lens = np.diff(cu)
    smax = int(lens.max())
    ids = np.zeros((n_seq, smax), dtype=np.int64)
    mask = np.zeros((n_seq, smax), dtype=np.int64)
    for s in range(n_seq):
        ids[s, :lens[s]] = tokens[cu[s]:cu[s + 1]]
        mask[s, :lens[s]] = 1
    with torch.no_grad():
        hs = m(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask)).last_hidden_state.numpy()
    out = np.zeros((n_seq, cfg.hidden))
    for s in range(n_seq):
        v = hs[s, 0] if cfg.pool == "cls" else hs[s, :lens[s]].mean(axis=0)
        if cfg.normalize:
            v = v / max(np.sqrt((v * v).sum()), 1e-12)
        out[s] = v
    return out


CASES = [
    ("tiny_cls", B.TINY, 5, 11, dict(mean_len=20, std_len=12, min_len=2, max_len=64)),
    ("tiny_mean", B.BertConfig(hidden=64, layers=2, heads=4, inter=128, vocab=120, max_pos=64, pool="mean"), 5, 12,
     dict(mean_len=20, std_len=12, min_len=2, max_len=64)),
    ("tiny_mean_nonorm", B.BertConfig(hidden=64, layers=2, heads=4, inter=128, vocab=120, max_pos=64, pool="mean",
                                      normalize=False), 3, 13, dict(mean_len=8, std_len=4, min_len=1, max_len=64)),
    # the real architectures at a few short sequences (outputs only: weights are re-seeded by the tests)
    ("minilm_l6", B.MINILM_L6, 3, 14, dict(mean_len=24, std_len=8, min_len=4, max_len=48)),
    ("bge_base", B.BGE_BASE, 2, 15, dict(mean_len=24, std_len=8, min_len=4, max_len=48)),
    # cfg4's 1024-d vectors come from this architecture (H1024 L24 A16 I4096)
    ("bge_large", B.BGE_LARGE, 2, 16, dict(mean_len=24, std_len=8, min_len=4, max_len=48)),
]


def main():
    out = {"generator": "oracle/gen_golden_bert.py", "transformers": __import__("transformers").__version__,
           "weights": "oracle.bert_encoder.init_weights(cfg, seed=7, bf16=True)", "cases": []}
    for name, cfg, n_seq, seed, kw in CASES:
        w = B.init_weights(cfg, seed=7, bf16=True)
        tok, cu = B.synth_batch(cfg, n_seq, seed, **kw)
        ref = hf_encode(cfg, w, tok, cu)
        mine = B.encode(cfg, w, tok, cu)
        err = float(np.abs(ref - mine).max())
        print(f"{name}: n_seq={n_seq} tokens={len(tok)} max|hf - oracle| = {err:.3e}")
        assert err < 1e-9, name
        out["cases"].append({"name": name, "cfg": cfg.__dict__, "n_seq": n_seq, "seed": seed, "batch": kw,
                             "tokens": tok.tolist(), "cu_seqlens": cu.tolist(),
                             "pooled": [[float(x) for x in row] for row in ref]})
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bert_ref.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
