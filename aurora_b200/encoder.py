"""Python handle over the CUDA text encoder (C ABI: aur_encoder_* in include/aurora_b200.h) and
the host-side mirror of the reference's embedding client.

Reference surface mirrored here: ``EmbeddingClient.embed / embed_batch / close`` and
``get_embedding_client()`` (server/services/correlation/embedding_client.py:20-84): ``None`` for
blank text or any failure, a list of floats otherwise.  The reference posts the text to a
t2v-transformers sidecar; here the forward pass runs in-process on the GPU.  Tokenisation is the
caller's (a ``tokenize(text) -> list[int]`` callable): a WordPiece tokenizer needs a vocabulary
file, which is deployment data, not code.
"""

from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass
from functools import lru_cache
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _native as N

logger = logging.getLogger(__name__)

POOL = {"cls": 0, "mean": 1}


@dataclass(frozen=True)
class EncoderConfig:
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    inter: int = 3072
    vocab: int = 30522
    max_pos: int = 512
    type_vocab: int = 2
    ln_eps: float = 1e-12
    pool: str = "cls"
    normalize: bool = True


BGE_BASE = EncoderConfig()
MINILM_L6 = EncoderConfig(hidden=384, layers=6, heads=12, inter=1536, pool="mean")   # head dim 32, zero-padded to 64 on the device
BGE_LARGE = EncoderConfig(hidden=1024, layers=24, heads=16, inter=4096)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_sequences(seqs: Sequence[Sequence[int]]):
    """list of token-id lists -> (tokens int32 [total], cu_seqlens int32 [n+1])."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    cu = np.zeros(len(seqs) + 1, dtype=np.int32)
    cu[1:] = np.cumsum(lens)
    tok = np.empty(int(cu[-1]), dtype=np.int32)
    for i, s in enumerate(seqs):
        tok[cu[i]:cu[i + 1]] = s
    return tok, cu


class Encoder:
    """BERT-family encoder resident on one GPU: packed token ids in, pooled vectors out."""

    def __init__(self, cfg: EncoderConfig, max_tokens: int = 32768, max_seqs: int = 2048, device: int = 0):
        self._lib = N.load()
        self.cfg, self.device = cfg, int(device)
        self.max_tokens, self.max_seqs = int(max_tokens), int(max_seqs)
        self._h = C.c_void_p()
        c = N.AurEncoderConfig(device=self.device, hidden=cfg.hidden, layers=cfg.layers, heads=cfg.heads, inter=cfg.inter,
                               vocab=cfg.vocab, max_pos=cfg.max_pos, type_vocab=cfg.type_vocab, pool=POOL[cfg.pool],
                               normalize=int(cfg.normalize), max_tokens=self.max_tokens, max_seqs=self.max_seqs,
                               ln_eps=cfg.ln_eps, reserved=0)
        N.check(self._lib.aur_encoder_open(C.byref(c), C.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.aur_encoder_close(self._h)
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

    # -------------------------------------------------------------- parameters
    def load_weights(self, weights: Dict[str, np.ndarray]) -> None:
        """Names and shapes as documented at aur_encoder_load (torch.nn.Linear layout)."""
        for name, w in weights.items():
            a = np.ascontiguousarray(w, dtype=np.float32)
            N.check(self._lib.aur_encoder_load(self._h, name.encode(), _ptr(a), a.size))

    # -------------------------------------------------------------- forward
    def _check_batch(self, tokens, cu):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        cu = np.ascontiguousarray(cu, dtype=np.int32)
        if cu.ndim != 1 or cu.size < 2 or tokens.ndim != 1 or tokens.size != int(cu[-1]):
            raise ValueError("tokens [total] and cu_seqlens [n_seq + 1] with cu[-1] == total are required")
        return tokens, cu

    def encode_packed(self, tokens: np.ndarray, cu_seqlens: np.ndarray, bf16: bool = False) -> np.ndarray:
        """[n_seq, hidden] float32 (or raw bf16 bits as uint16 when ``bf16``)."""
        tokens, cu = self._check_batch(tokens, cu_seqlens)
        n = cu.size - 1
        out = np.empty((n, self.cfg.hidden), dtype=np.uint16 if bf16 else np.float32)
        N.check(self._lib.aur_encode(self._h, _ptr(tokens), _ptr(cu), n, None if bf16 else _ptr(out),
                                     _ptr(out) if bf16 else None))
        return out

    def encode(self, seqs: Sequence[Sequence[int]]) -> np.ndarray:
        """Token-id lists of any count: split into calls that fit the workspace."""
        out = np.empty((len(seqs), self.cfg.hidden), dtype=np.float32)
        i = 0
        while i < len(seqs):
            j, toks = i, 0
            while j < len(seqs) and j - i < self.max_seqs and toks + len(seqs[j]) <= self.max_tokens:
                toks += len(seqs[j]); j += 1
            if j == i:
                raise ValueError(f"sequence {i} has {len(seqs[i])} tokens: more than max_tokens={self.max_tokens}")
            tok, cu = pack_sequences(seqs[i:j])
            out[i:j] = self.encode_packed(tok, cu)
            i = j
        return out

    def encode_append(self, index, tokens: np.ndarray, cu_seqlens: np.ndarray, ids: np.ndarray,
                      user_codes: Optional[np.ndarray] = None, org_codes: Optional[np.ndarray] = None) -> None:
        """Fused ingest: encode and append to ``index`` (engine.Index) without leaving the device."""
        tokens, cu = self._check_batch(tokens, cu_seqlens)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        u = None if user_codes is None else np.ascontiguousarray(user_codes, dtype=np.int32)
        o = None if org_codes is None else np.ascontiguousarray(org_codes, dtype=np.int32)
        N.check(self._lib.aur_encode_append(self._h, index._h, _ptr(tokens), _ptr(cu), cu.size - 1, _ptr(ids), _ptr(u), _ptr(o)))

    def stats(self) -> dict:
        st = N.AurEncoderStats()
        N.check(self._lib.aur_encoder_get_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_}

    def hidden_states(self) -> np.ndarray:
        """Final hidden states [tokens, hidden] (bf16 bits) of the last call (test hook)."""
        t = self.stats()["tokens"]
        out = np.empty((t, self.cfg.hidden), dtype=np.uint16)
        N.check(self._lib.aur_debug_encoder_hidden(self._h, _ptr(out), out.size))
        return out


class TextEncoder:
    """Adapter for aurora_b200.retriever.KnowledgeBase: ``dim`` + ``encode(texts)`` over a tokenizer and the CUDA
    ``Encoder``; ``encode_append`` is the fused ingest path (vectors never leave the device between the forward
    pass and the shard append).

    ``tokenizer``: an ``aurora_b200.wordpiece.NativeTokenizer`` (C++, multi-threaded -- then whole batches go through
    ONE C call, ``aur_encode_text_append``, with no Python per text) or any ``tokenize(text) -> list[int]`` callable.

    Texts longer than the position table (the reference's chunker can emit a 4 000-character chunk,
    document_processor.py:266-267) are not cut off: they are split into windows of ``max_pos`` tokens, each window is
    encoded, and the text's vector is the L2-normalised mean of its windows -- the same idea as the t2v sidecar's
    sentence batching + averaging (that container's exact splitting rule is not in the reference tree)."""

    def __init__(self, encoder: Encoder, tokenizer, max_len: Optional[int] = None):
        self._enc, self._tok = encoder, tokenizer
        self._native = hasattr(tokenizer, "encode_packed")
        self.dim = encoder.cfg.hidden
        self.max_len = int(max_len or encoder.cfg.max_pos)

    # ------------------------------------------------------------------ tokenisation
    def _tokenize(self, texts: Sequence[str], max_len: int):
        if self._native:
            return self._tok.encode_packed(list(texts), max_len)
        seqs = []
        for t in texts:
            ids = list(self._tok(t))
            seqs.append(ids if len(ids) <= max_len else ids[: max_len - 1] + ids[-1:])
        return pack_sequences(seqs)

    def _over_long(self, texts: Sequence[str], lens: np.ndarray) -> List[int]:
        """Indices of the texts whose full tokenisation exceeds the position table.  Only texts that came back at
        exactly ``max_len`` ids can have been truncated; those are re-measured without the limit -- with the native
        tokenizer in ONE C call for all of them."""
        cand = [int(i) for i in np.nonzero(np.asarray(lens) >= self.max_len)[0]]
        if not cand:
            return []
        if self._native and hasattr(self._tok, "lengths"):
            full = self._tok.lengths([texts[i] for i in cand])
            return [i for i, n in zip(cand, full) if n > self.max_len]
        return [i for i in cand if len(self._windows(texts[i])) > 1]

    def _windows(self, text: str):
        """Token windows of one over-long text: [CLS] body[i : i + max_len - 2] [SEP]."""
        tok, cu = self._tokenize([text], 1 << 20) if self._native else pack_sequences([list(self._tok(text))])
        ids = tok[cu[0]:cu[1]]
        cls, sep, body = ids[0], ids[-1], ids[1:-1]
        step = self.max_len - 2
        return [np.concatenate(([cls], body[i:i + step], [sep])).astype(np.int32) for i in range(0, max(len(body), 1), step)]

    def _encode_packed_any(self, tok: np.ndarray, cu: np.ndarray) -> np.ndarray:
        out = np.empty((len(cu) - 1, self.dim), dtype=np.float32)
        i = 0
        while i < len(cu) - 1:
            j = i
            while j < len(cu) - 1 and j - i < self._enc.max_seqs and cu[j + 1] - cu[i] <= self._enc.max_tokens:
                j += 1
            if j == i:
                raise ValueError(f"text {i} tokenizes to {cu[i + 1] - cu[i]} tokens: more than max_tokens={self._enc.max_tokens}")
            out[i:j] = self._enc.encode_packed(tok[cu[i]:cu[j]], (cu[i:j + 1] - cu[i]).astype(np.int32))
            i = j
        return out

    def _encode_long(self, text: str) -> np.ndarray:
        return self._encode_long_many([text])[0]

    def _encode_long_many(self, texts: Sequence[str]) -> np.ndarray:
        """Vectors of over-long texts: all their windows go through the encoder as ONE packed batch, then each text's
        windows are averaged (and re-normalised when the encoder normalises)."""
        wins, owner = [], []
        for t, text in enumerate(texts):
            w = self._windows(text)
            wins += w
            owner += [t] * len(w)
        tok, cu = pack_sequences(wins)
        vecs = self._encode_packed_any(tok, cu)
        owner = np.asarray(owner)
        out = np.empty((len(texts), self.dim), dtype=np.float32)
        for t in range(len(texts)):
            v = vecs[owner == t].mean(axis=0)
            n = float(np.linalg.norm(v))
            out[t] = v / n if (self._enc.cfg.normalize and n > 0) else v
        return out

    # ------------------------------------------------------------------ API used by the retriever
    def encode(self, texts: Sequence[str]) -> np.ndarray:
        tok, cu = self._tokenize(texts, self.max_len)
        out = self._encode_packed_any(tok, cu)
        long_ones = self._over_long(texts, np.diff(cu))               # truncated above: windows, averaged
        if long_ones:
            out[long_ones] = self._encode_long_many([texts[i] for i in long_ones])
        return out

    def encode_append(self, index, texts: Sequence[str], ids: np.ndarray, user_codes=None, org_codes=None) -> None:
        """Tokenise once (one multi-threaded C call with the native tokenizer), then encoder forward + shard append per
        workspace-sized batch; nothing runs per text in Python."""
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        tok, cu = self._tokenize(texts, self.max_len)
        lens = np.diff(cu)
        long_ones = self._over_long(texts, lens)
        u = None if user_codes is None else np.ascontiguousarray(user_codes, dtype=np.int32)
        o = None if org_codes is None else np.ascontiguousarray(org_codes, dtype=np.int32)
        keep = np.ones(len(texts), dtype=bool)
        keep[long_ones] = False
        sel = np.nonzero(keep)[0]
        i = 0
        while i < len(sel):
            j, toks = i, 0
            while j < len(sel) and j - i < self._enc.max_seqs and toks + lens[sel[j]] <= self._enc.max_tokens:
                toks += int(lens[sel[j]]); j += 1
            if j == i:
                raise ValueError(f"text {sel[i]} tokenizes to {lens[sel[i]]} tokens: more than max_tokens={self._enc.max_tokens}")
            part = sel[i:j]
            if len(part) == part[-1] - part[0] + 1:            # contiguous run: slice the packed buffer
                t = tok[cu[part[0]]:cu[part[-1] + 1]]
                c = (cu[part[0]:part[-1] + 2] - cu[part[0]]).astype(np.int32)
            else:
                t, c = pack_sequences([tok[cu[x]:cu[x + 1]] for x in part])
            self._enc.encode_append(index, t, c, ids[part], None if u is None else u[part], None if o is None else o[part])
            i = j
        if long_ones:                                                 # averaged windows: the vectors are formed on the host
            v = self._encode_long_many([texts[i] for i in long_ones])
            index.add(v, ids[long_ones], None if u is None else u[long_ones], None if o is None else o[long_ones])

    def _append_texts_native(self, index, texts, ids, u, o) -> None:
        """The same ingest as ONE C call (aur_encode_text_append: tokenise + batches + append inside the library), for
        callers that know their texts fit the position table (over-long texts would be truncated there)."""
        blob, offs = self._tok._pack(list(texts))
        N.check(self._enc._lib.aur_encode_text_append(self._enc._h, self._tok._h, index._h, blob, _ptr(offs), len(texts), self.max_len,
                                                      self._enc.max_tokens, self._enc.max_seqs, _ptr(ids), _ptr(u), _ptr(o), 0))


# ----------------------------------------------------------------------------- reference mirror
class EmbeddingClient:
    """Drop-in for server/services/correlation/embedding_client.py:20-78 backed by ``Encoder``."""

    def __init__(self, encoder: Encoder, tokenize: Callable[[str], List[int]]):
        self._enc, self._tok = encoder, tokenize

    def embed(self, text: str) -> Optional[List[float]]:
        if not text or not text.strip():                      # embedding_client.py:48-49
            return None
        try:
            return self._enc.encode([self._tok(text)])[0].tolist()
        except Exception as e:                                  # embedding_client.py:66-70: any failure -> None
            logger.warning("[EmbeddingClient] embed failed: %s", e)
            return None

    def embed_batch(self, texts: List[str]) -> List[Optional[List[float]]]:
        """One batched forward instead of the reference's per-text loop (embedding_client.py:72-74);
        blank texts and failures still map to None per element."""
        idx = [i for i, t in enumerate(texts) if t and t.strip()]
        out: List[Optional[List[float]]] = [None] * len(texts)
        if not idx:
            return out
        try:
            vecs = self._enc.encode([self._tok(texts[i]) for i in idx])
            for j, i in enumerate(idx):
                out[i] = vecs[j].tolist()
        except Exception as e:
            logger.warning("[EmbeddingClient] embed_batch failed: %s", e)
        return out

    def close(self) -> None:                                    # embedding_client.py:76-78
        self._enc.close()


_factory: Optional[Callable[[], EmbeddingClient]] = None


def configure_embedding_client(factory: Callable[[], EmbeddingClient]) -> None:
    """Deployment hook: how to build the process-wide client (weights, tokenizer, device)."""
    global _factory
    _factory = factory
    get_embedding_client.cache_clear()


@lru_cache(maxsize=1)
def get_embedding_client() -> EmbeddingClient:                 # embedding_client.py:81-84
    if _factory is None:
        raise RuntimeError("call aurora_b200.encoder.configure_embedding_client(factory) first")
    return _factory()


# ----------------------------------------------------------------------------- checkpoint loading
def read_safetensors(path: str) -> Dict[str, np.ndarray]:
    """Minimal safetensors reader (8-byte little-endian header length, JSON header, raw tensors):
    F32 / F16 / BF16 tensors come back as float32 arrays.  No third-party dependency."""
    import json
    import struct

    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen))
        base = 8 + hlen
        for name, info in header.items():
            if name == "__metadata__":
                continue
            lo, hi = info["data_offsets"]
            f.seek(base + lo)
            raw = f.read(hi - lo)
            dt = info["dtype"]
            if dt == "F32":
                a = np.frombuffer(raw, dtype=np.float32)
            elif dt == "F16":
                a = np.frombuffer(raw, dtype=np.float16).astype(np.float32)
            elif dt == "BF16":
                a = (np.frombuffer(raw, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)
            else:
                continue                                   # integer buffers (position_ids) are not parameters
            out[name] = a.reshape(info["shape"]).copy()
    return out


def from_hf_bert(state: Dict[str, np.ndarray], cfg: EncoderConfig) -> Dict[str, np.ndarray]:
    """HuggingFace BertModel parameter names (with or without a ``bert.`` prefix) -> the names
    ``aur_encoder_load`` expects; query / key / value matrices are stacked into ``wqkv``."""
    def g(name: str) -> np.ndarray:
        for prefix in ("", "bert."):
            if prefix + name in state:
                return np.asarray(state[prefix + name], dtype=np.float32)
        raise KeyError(name)

    w = {"word_emb": g("embeddings.word_embeddings.weight"), "pos_emb": g("embeddings.position_embeddings.weight"),
         "type_emb": g("embeddings.token_type_embeddings.weight"), "emb_ln_g": g("embeddings.LayerNorm.weight"),
         "emb_ln_b": g("embeddings.LayerNorm.bias")}
    for l in range(cfg.layers):
        p, q = f"l{l}.", f"encoder.layer.{l}."
        w[p + "wqkv"] = np.concatenate([g(q + f"attention.self.{n}.weight") for n in ("query", "key", "value")], axis=0)
        w[p + "bqkv"] = np.concatenate([g(q + f"attention.self.{n}.bias") for n in ("query", "key", "value")], axis=0)
        w[p + "wo"], w[p + "bo"] = g(q + "attention.output.dense.weight"), g(q + "attention.output.dense.bias")
        w[p + "ln1_g"], w[p + "ln1_b"] = g(q + "attention.output.LayerNorm.weight"), g(q + "attention.output.LayerNorm.bias")
        w[p + "wi"], w[p + "bi"] = g(q + "intermediate.dense.weight"), g(q + "intermediate.dense.bias")
        w[p + "wo2"], w[p + "bo2"] = g(q + "output.dense.weight"), g(q + "output.dense.bias")
        w[p + "ln2_g"], w[p + "ln2_b"] = g(q + "output.LayerNorm.weight"), g(q + "output.LayerNorm.bias")
    return w
