"""Checkpoint path of the encoder: safetensors file -> HF BertModel names -> aur_encoder_load names.
CPU only (no device needed to read a file); the loaded dict is compared with the oracle's own weights."""

import json
import struct

import numpy as np

from aurora_b200.encoder import EncoderConfig, from_hf_bert, read_safetensors
from oracle import bert_encoder as B
from oracle.cosine_topk import f32_to_bf16_bits


def _write_safetensors(path, tensors):
    header, blobs, off = {}, [], 0
    for name, (arr, dt) in tensors.items():
        raw = {"F32": lambda a: a.astype("<f4").tobytes(), "F16": lambda a: a.astype("<f2").tobytes(),
               "BF16": lambda a: f32_to_bf16_bits(a.astype(np.float32)).astype("<u2").tobytes()}[dt](arr)
        header[name] = {"dtype": dt, "shape": list(arr.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw); off += len(raw)
    header["__metadata__"] = {"format": "pt"}
    hj = json.dumps(header).encode()
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj))); f.write(hj)
        for b in blobs:
            f.write(b)


def test_safetensors_to_encoder_weights(tmp_path):
    cfg_o = B.TINY
    w = B.init_weights(cfg_o, seed=7, bf16=True)
    h = cfg_o.hidden
    hf = {"bert.embeddings.word_embeddings.weight": (w["word_emb"], "BF16"),
          "bert.embeddings.position_embeddings.weight": (w["pos_emb"], "F32"),
          "bert.embeddings.token_type_embeddings.weight": (w["type_emb"], "F16"),
          "bert.embeddings.LayerNorm.weight": (w["emb_ln_g"], "F32"), "bert.embeddings.LayerNorm.bias": (w["emb_ln_b"], "F32")}
    for l in range(cfg_o.layers):
        p, q = f"l{l}.", f"bert.encoder.layer.{l}."
        for i, n in enumerate(("query", "key", "value")):
            hf[q + f"attention.self.{n}.weight"] = (w[p + "wqkv"][i * h:(i + 1) * h], "BF16")
            hf[q + f"attention.self.{n}.bias"] = (w[p + "bqkv"][i * h:(i + 1) * h], "F32")
        for a, b in (("attention.output.dense.weight", "wo"), ("attention.output.dense.bias", "bo"),
                     ("attention.output.LayerNorm.weight", "ln1_g"), ("attention.output.LayerNorm.bias", "ln1_b"),
                     ("intermediate.dense.weight", "wi"), ("intermediate.dense.bias", "bi"), ("output.dense.weight", "wo2"),
                     ("output.dense.bias", "bo2"), ("output.LayerNorm.weight", "ln2_g"), ("output.LayerNorm.bias", "ln2_b")):
            hf[q + a] = (w[p + b], "F32")
    path = str(tmp_path / "model.safetensors")
    _write_safetensors(path, hf)
    cfg = EncoderConfig(hidden=h, layers=cfg_o.layers, heads=cfg_o.heads, inter=cfg_o.inter, vocab=cfg_o.vocab, max_pos=cfg_o.max_pos)
    got = from_hf_bert(read_safetensors(path), cfg)
    assert set(got) == set(B.weight_names(cfg_o))
    for name in got:
        tol = 1e-3 if name == "type_emb" else 0.0          # the F16 tensor is not bf16-exact
        np.testing.assert_allclose(got[name], w[name], rtol=0, atol=tol, err_msg=name)


def test_pack_sequences_and_presets():
    from aurora_b200.encoder import BGE_BASE, BGE_LARGE, MINILM_L6, pack_sequences

    tok, cu = pack_sequences([[101, 5, 102], [101, 102], [101, 7, 8, 9, 102]])
    assert tok.dtype == np.int32 and cu.dtype == np.int32
    assert tok.tolist() == [101, 5, 102, 101, 102, 101, 7, 8, 9, 102] and cu.tolist() == [0, 3, 5, 10]
    assert (BGE_BASE.hidden, BGE_BASE.heads, BGE_BASE.layers, BGE_BASE.pool) == (768, 12, 12, "cls")
    assert (MINILM_L6.hidden, MINILM_L6.heads, MINILM_L6.layers, MINILM_L6.pool) == (384, 12, 6, "mean")   # head dim 32
    assert BGE_LARGE.hidden // BGE_LARGE.heads == 64
    # the oracle's presets describe the same architectures
    for mine, theirs in ((BGE_BASE, B.BGE_BASE), (MINILM_L6, B.MINILM_L6), (BGE_LARGE, B.BGE_LARGE)):
        assert (mine.hidden, mine.layers, mine.heads, mine.inter, mine.vocab, mine.pool) == \
               (theirs.hidden, theirs.layers, theirs.heads, theirs.inter, theirs.vocab, theirs.pool)
