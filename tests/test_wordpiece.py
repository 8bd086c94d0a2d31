"""aurora_b200.wordpiece against transformers.BertTokenizer (the tokenizer the reference's t2v sidecar
applies before its BERT forward) on a synthetic uncased vocabulary -- no vocabulary file ships with
this repo or is downloadable here."""

import os

import pytest

from aurora_b200.wordpiece import WordPieceTokenizer, basic_tokenize, load_vocab

TEXTS = [
    "Restart the payment-service when p99 latency > 2s (see runbook #42).",
    "Kafka consumer-lag alert: partition 7 is 1,234,567 msgs behind!",
    "Café déjà-vu naïve coöperate — unicode dashes… and “quotes”",
    "数据库连接池耗尽 database pool exhausted",
    "   multiple   spaces\tand\nnewlines  ",
    "unknownword zzzzqqq xylophone",
    "a" * 120 + " short",
    "",
]


def _vocab(tmp_path):
    pieces = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    words = set()
    for t in TEXTS:
        for w in basic_tokenize(t):
            words.add(w)
    for w in sorted(words):
        if w in ("zzzzqqq", "xylophone") or len(w) > 100:
            continue                                    # left out: must become [UNK]
        if len(w) > 4:                                  # split long words into a head and ## tails
            pieces += [w[:3], "##" + w[3:5], "##" + w[5:]] if len(w) > 5 else [w[:3], "##" + w[3:]]
        else:
            pieces.append(w)
    pieces += ["##s", "##ing", "re", "##start"]
    seen, uniq = set(), []
    for p in pieces:
        if p and p != "##" and p not in seen:
            seen.add(p); uniq.append(p)
    path = os.path.join(tmp_path, "vocab.txt")
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(uniq) + "\n")
    return path


def test_matches_hf_bert_tokenizer(tmp_path):
    transformers = pytest.importorskip("transformers")
    path = _vocab(str(tmp_path))
    try:
        hf = transformers.BertTokenizer(vocab=load_vocab(path), do_lower_case=True)
    except Exception as e:                              # pragma: no cover - tokenizer backend unavailable
        pytest.skip(f"BertTokenizer unavailable: {e}")
    mine = WordPieceTokenizer(load_vocab(path))
    for t in TEXTS:
        want = hf(t, truncation=True, max_length=32)["input_ids"]
        assert mine.encode(t, max_len=32) == want, t
        assert mine.tokenize_ids(t) == hf.convert_tokens_to_ids(hf.tokenize(t)), t


def test_unknown_and_truncation(tmp_path):
    mine = WordPieceTokenizer(load_vocab(_vocab(str(tmp_path))))
    ids = mine.encode("zzzzqqq restart")
    assert ids[0] == mine.cls_id and ids[-1] == mine.sep_id and ids[1] == mine.unk_id
    assert len(mine.encode("restart " * 600, max_len=512)) == 512
    assert mine.encode("") == [mine.cls_id, mine.sep_id]
