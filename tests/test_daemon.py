"""Engine daemon + client shim: the reference's module API over a Unix socket (CPU, test doubles)."""

import threading

import pytest

from aurora_b200 import retriever as R
from aurora_b200.daemon import Client, serve
from tests.doubles import HashEmbedder, OracleIndex


@pytest.fixture()
def daemon(tmp_path):
    R.configure(encoder=HashEmbedder(64), capacity=1024, index_factory=lambda dim, cap: OracleIndex(dim, cap))
    path = str(tmp_path / "kb.sock")
    srv = serve(path, background=True)
    yield path
    srv.shutdown(); srv.server_close()
    R.configure(factory=lambda: (_ for _ in ()).throw(RuntimeError("backend down")))


def _chunks(*texts):
    return [{"content": t, "heading_context": "", "chunk_index": i} for i, t in enumerate(texts)]


def test_client_roundtrip_and_concurrency(daemon):
    kb = Client(daemon)
    assert kb.health()["ready"]
    assert kb.insert_chunks("u", "d", "f.md", _chunks("redis failover steps", "postgres vacuum", "kafka lag")) == 3
    res = kb.search_knowledge_base("u", "redis failover", limit=2)
    assert res[0]["content"] == "redis failover steps" and set(res[0]) == {"content", "heading_context", "source_filename",
                                                                           "document_id", "chunk_index", "score"}
    assert kb.get_document_chunk_count("u", "d") == 3
    errs = []

    def worker(i):
        try:
            c = Client(daemon)
            for _ in range(20):
                assert c.search_knowledge_base("u", "kafka lag", limit=1)[0]["chunk_index"] == 2
        except Exception as e:      # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(8)]   # 2 gunicorn workers x 4 threads
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errs
    assert kb.delete_document_chunks("u", "d") == 3 and kb.search_knowledge_base("u", "redis") == []


def test_error_conventions_without_a_daemon(tmp_path):
    kb = Client(str(tmp_path / "nobody.sock"))
    assert kb.health()["ready"] is False
    assert kb.search_knowledge_base("u", "q") == []
    assert kb.delete_document_chunks("u", "d") == -1 and kb.delete_user_chunks("u") == -1
    assert kb.get_document_chunk_count("u", "d") == 0 and kb.delete_discovery_chunks("o") == 0
    assert kb.insert_chunks("u", "d", "f", []) == 0
    with pytest.raises(Exception):
        kb.insert_chunks("u", "d", "f", _chunks("x"))


def test_backend_failure_is_reported_not_swallowed_for_insert(daemon):
    R.configure(factory=lambda: (_ for _ in ()).throw(RuntimeError("backend down")))
    kb = Client(daemon)
    with pytest.raises(RuntimeError):
        kb.insert_chunks("u", "d", "f", _chunks("x"))          # Celery retries (weaviate_client.py:210-212)
    assert kb.search_knowledge_base("u", "q") == []


def test_bootstrap_requires_its_variables(monkeypatch):
    from aurora_b200 import bootstrap

    monkeypatch.delenv("AURORA_B200_ENCODER_WEIGHTS", raising=False)
    monkeypatch.delenv("AURORA_B200_VOCAB", raising=False)
    with pytest.raises(RuntimeError, match="AURORA_B200_ENCODER_WEIGHTS"):
        bootstrap.configure_from_env()
    with pytest.raises(ValueError):
        bootstrap._model_config("gpt-2")
    assert bootstrap._model_config("minilm-l6").hidden == 384
