"""Test doubles (tests only): a deterministic text embedder and a CPU stand-in for the GPU
index so the retriever's host logic can be exercised without a device.  The reference's own
suite does the same thing one level up: it replaces `weaviate` with MagicMock
(server/tests/conftest.py:41-61)."""

import hashlib
import re

import numpy as np

from oracle import cosine_topk as O


class HashEmbedder:
    """Bag-of-words hashing embedder: similar texts -> similar vectors.  Deterministic."""

    def __init__(self, dim: int = 64):
        self.dim = dim
        self.calls = 0

    def encode(self, texts):
        self.calls += 1
        out = np.zeros((len(texts), self.dim), dtype=np.float32)
        for i, t in enumerate(texts):
            for tok in re.findall(r"[a-z0-9]+", t.lower()):
                h = int.from_bytes(hashlib.sha1(tok.encode()).digest()[:8], "little")
                out[i, h % self.dim] += 1.0
                out[i, (h >> 20) % self.dim] += 0.5
        return out


class OracleIndex:
    """Same surface as aurora_b200.engine.Index, backed by the CPU oracle (bf16 store)."""

    def __init__(self, dim, capacity):
        self.dim, self.capacity = dim, capacity
        self.rows = np.zeros((0, dim), dtype=np.float32)
        self.ids = np.zeros(0, dtype=np.int64)
        self.user = np.zeros(0, dtype=np.int32)
        self.org = np.zeros(0, dtype=np.int32)
        self.live = np.zeros(0, dtype=bool)

    def add(self, rows, ids, user_codes=None, org_codes=None):
        rows = O.round_to_bf16(np.asarray(rows, dtype=np.float32))
        n = rows.shape[0]
        if len(self.ids) + n > self.capacity:
            raise RuntimeError("shard full")
        for i in ids:                                   # upsert: old row becomes a tombstone
            self.live[self.ids == i] = False
        self.rows = np.concatenate([self.rows, rows])
        self.ids = np.concatenate([self.ids, np.asarray(ids, dtype=np.int64)])
        self.user = np.concatenate([self.user, np.zeros(n, np.int32) if user_codes is None else user_codes])
        self.org = np.concatenate([self.org, np.full(n, -1, np.int32) if org_codes is None else org_codes])
        self.live = np.concatenate([self.live, np.ones(n, dtype=bool)])

    def remove(self, ids):
        hit = np.isin(self.ids, np.asarray(ids)) & self.live
        self.live[hit] = False
        return int(hit.sum())

    def search(self, queries, k, q_user=None, q_org=None):
        q = O.round_to_bf16(np.asarray(queries, dtype=np.float32))
        return O.cosine_topk(q, self.rows, k, ids=self.ids, live=self.live, row_user=self.user, row_org=self.org,
                             q_user=q_user, q_org=q_org)

    def search_subset(self, queries, k, allow_ids):
        q = O.round_to_bf16(np.asarray(queries, dtype=np.float32))
        live = self.live & np.isin(self.ids, np.asarray(allow_ids, dtype=np.int64))
        return O.cosine_topk(q, self.rows, k, ids=self.ids, live=live)

    def save(self, path):
        np.savez(path, rows=self.rows[self.live], ids=self.ids[self.live], user=self.user[self.live], org=self.org[self.live],
                 dim=self.dim, capacity=self.capacity)

    @classmethod
    def load(cls, path, capacity=None):
        z = np.load(path if path.endswith(".npz") else path + ".npz")
        ix = cls(int(z["dim"]), int(capacity or z["capacity"]))
        if len(z["ids"]):
            ix.add(z["rows"], z["ids"], z["user"], z["org"])
        return ix

    def export(self):
        return self.rows, self.ids, self.user, self.org, self.live

    def stats(self):
        return {"rows": int(len(self.ids)), "live": int(self.live.sum()), "searches": 0, "last_kernel": 0}

    def compact(self):
        dead = int((~self.live).sum())
        keep = self.live
        self.rows, self.ids, self.user, self.org = self.rows[keep], self.ids[keep], self.user[keep], self.org[keep]
        self.live = np.ones(len(self.ids), dtype=bool)
        return dead

    def sync(self):
        pass

    def close(self):
        pass
