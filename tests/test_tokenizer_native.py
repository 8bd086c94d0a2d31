"""The C++ WordPiece tokenizer behind the C ABI (csrc/tokenizer.cpp) against transformers.BertTokenizer -- the
tokenizer the reference's t2v sidecar applies -- id for id, on runbook-style text, Unicode edge cases and seeded
random strings; plus the pure-Python restatement on ordinary text.  CPU only (no GPU call)."""

import random

import numpy as np
import pytest

from aurora_b200.wordpiece import NativeTokenizer, WordPieceTokenizer

transformers = pytest.importorskip("transformers")
U = lambda *cps: "".join(chr(c) for c in cps)      # noqa: E731

TEXTS = [
    "Restart the payment-service when p99 latency > 2s (see runbook #42).",
    "Kafka consumer-lag alert: partition 7 is 1,234,567 msgs behind!",
    "Caf" + U(0xe9) + " d" + U(0xe9) + "j" + U(0xe0) + "-vu na" + U(0xef) + "ve co" + U(0xf6) + "perate " + U(0x2014) + " unicode dashes" + U(0x2026) + " and " + U(0x201c) + "quotes" + U(0x201d),
    U(0x6570, 0x636e, 0x5e93, 0x8fde, 0x63a5, 0x6c60, 0x8017, 0x5c3d) + " database pool exhausted",
    "   multiple   spaces\tand\nnewlines  ",
    "unknownword zzzzqqq xylophone",
    "a" * 120 + " short",
    "",
    "MiXeD CaSe And ALLCAPS and camelCaseWord",
]
EDGE = [U(0x39f, 0x394, 0x3a5, 0x3a3, 0x3a3, 0x395, 0x3a5, 0x3a3), U(0x130), U(0x1e9e), U(0x1c5), U(0xc5), U(0x61, 0xad, 0x62), U(0x78, 0x200b, 0x79),
        U(0xfeff, 0x62), U(0x61, 0x20, 0x62, 0xfffd, 0x63), U(0xff46, 0xff55), U(0x2460), U(0xd55c, 0xae00), U(0x61, 0xa0, 0x62),
        U(0x61, 0x2028, 0x62), U(0xac00), U(0x61, 0x7f, 0x62), U(0x61, 0x85, 0x62), U(0x61, 0x378, 0x62), U(0x65, 0x301), U(0x1e69),
        U(0x61, 0x20dd, 0x62), U(0x61, 0x903, 0x62), U(0x61, 0x0, 0x62), U(0x61, 0xe000, 0x62), U(0x61, 0x1f600, 0x62), U(0x61, 0xb7, 0x62),
        U(0x61, 0x2d, 0x62), U(0x61, 0x60, 0x62), U(0x61, 0x5e, 0x62), U(0x61, 0x24, 0x62), U(0x61, 0xa2, 0x62), U(0x61, 0x3001, 0x62),
        U(0x4e2d, 0x6587), U(0x1e9b, 0x323), U(0xf900, 0x61), U(0x2f800), U(0x301, 0x61), U(0x1100, 0x1161), U(0x3b1, 0x345, 0x301)]


def _fuzz(n=300, seed=5):
    rng = random.Random(seed)
    pools = [(0x20, 0x7e), (0xa0, 0x24f), (0x370, 0x3ff), (0x400, 0x4ff), (0x300, 0x36f), (0x2000, 0x206f), (0x3000, 0x303f),
             (0x4e00, 0x4e80), (0xac00, 0xac80), (0x1f600, 0x1f640), (0x0, 0x1f), (0xff00, 0xff60), (0x1e00, 0x1eff), (0x900, 0x97f)]
    out = []
    for _ in range(n):
        s = []
        for _ in range(rng.randint(1, 40)):
            lo, hi = rng.choice(pools if rng.random() < 0.5 else pools[:2])
            cp = rng.randint(lo, hi)
            if 0xd800 <= cp <= 0xdfff:
                continue
            s.append(chr(cp))
        out.append("".join(s))
    return out


def _char_vocab(hf_norm, texts):
    """[specials] + every normalised character as a word-initial and as a ## piece + a few multi-character pieces."""
    pieces = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    chars = set()
    for t in texts:
        for c in hf_norm.normalize_str(t):
            if not c.isspace():
                chars.add(c)
    for c in sorted(chars):
        pieces += [c, "##" + c]
    pieces += ["restart", "the", "pay", "##ment", "service", "latency", "run", "##book", "kafka", "consumer", "lag", "alert", "data", "##base",
               "pool", "multiple", "spaces", "and", "new", "##lines", "short", "un", "##known", "##word", "mixed", "case", "all", "##caps", "camel"]
    seen, uniq = set(), []
    for p in pieces:
        if p not in seen:
            seen.add(p); uniq.append(p)
    return {p: i for i, p in enumerate(uniq)}


@pytest.fixture(scope="module")
def pair():
    all_texts = TEXTS + EDGE + _fuzz()
    boot = transformers.BertTokenizer(vocab={t: i for i, t in enumerate(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"])}, do_lower_case=True)
    vocab = _char_vocab(boot.backend_tokenizer.normalizer, all_texts)
    hf = transformers.BertTokenizer(vocab=vocab, do_lower_case=True)
    return hf, NativeTokenizer(vocab, lower=True), vocab, all_texts


def test_native_matches_hf_on_text_edge_cases_and_fuzz(pair):
    hf, nat, vocab, all_texts = pair
    assert (nat.vocab_size, nat.unk_id, nat.cls_id, nat.sep_id) == (len(vocab), vocab["[UNK]"], vocab["[CLS]"], vocab["[SEP]"])
    got = nat.encode_batch(all_texts, max_len=64)
    for t, g in zip(all_texts, got):
        want = hf(t, truncation=True, max_length=64)["input_ids"]
        assert g == want, [hex(ord(c)) for c in t[:40]]


def test_native_truncation_packing_and_threads(pair):
    hf, nat, vocab, _ = pair
    texts = ["restart the payment service " * 300, "", "kafka consumer lag alert", "a" * 101 + " b"] * 40
    tok1, cu1 = nat.encode_packed(texts, max_len=512, threads=1)
    tok8, cu8 = nat.encode_packed(texts, max_len=512, threads=8)
    assert np.array_equal(tok1, tok8) and np.array_equal(cu1, cu8)
    lens = np.diff(cu1)
    assert lens[0] == 512 and lens[1] == 2 and tok1[cu1[1]] == nat.cls_id and tok1[cu1[2] - 1] == nat.sep_id
    assert nat.encode("a" * 101 + " b")[1] == nat.unk_id                         # > 100 characters -> [UNK]
    assert nat.encode(texts[0], max_len=16) == hf(texts[0], truncation=True, max_length=16)["input_ids"]
    with pytest.raises(Exception):
        nat.encode_packed(["x"], max_len=1)
    # count mode (tokens_out = NULL): the lengths alone, with and without the limit
    assert np.array_equal(nat.lengths(texts, max_len=512), lens)
    full = nat.lengths(texts)
    assert full[0] == len(hf(texts[0])["input_ids"]) > 512 and np.array_equal(full[1:4], lens[1:4])


def test_cased_vocabulary_keeps_case_and_accents():
    vocab = {p: i for i, p in enumerate(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "Caf", "##" + U(0xe9), "caf", "##e", "A", "a"])}
    hf = transformers.BertTokenizer(vocab=vocab, do_lower_case=False)
    nat = NativeTokenizer(vocab, lower=False)
    for t in ["Caf" + U(0xe9), "caf" + U(0xe9), "A a", "CAF"]:
        assert nat.encode(t) == hf(t)["input_ids"], t


def test_python_restatement_agrees_on_ordinary_text(pair):
    _, nat, vocab, _ = pair
    py = WordPieceTokenizer(vocab)
    for t in TEXTS:
        assert py.encode(t, max_len=64) == nat.encode(t, max_len=64), t
