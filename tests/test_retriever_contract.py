"""Drop-in contract of aurora_b200.retriever vs server/routes/knowledge_base/weaviate_client.py:
same names, keyword arguments, return shapes and error conventions.  CPU only (the vector
index is a test double; the GPU-backed run of the same contract is in test_gpu_retriever.py)."""

import os
import inspect

import pytest

from aurora_b200 import retriever as R
from aurora_b200.filters import Filter, HybridFusion
from tests.doubles import HashEmbedder, OracleIndex


@pytest.fixture()
def kb():
    R.configure(encoder=HashEmbedder(64), capacity=4096, index_factory=lambda dim, cap: OracleIndex(dim, cap))
    yield R
    R.configure(factory=lambda: (_ for _ in ()).throw(RuntimeError("backend down")))


def _chunks(*texts):
    return [{"content": t, "heading_context": f"Doc > S{i}", "chunk_index": i} for i, t in enumerate(texts)]


def test_signatures_match_the_reference_module():
    # weaviate_client.py:136-142, :215-222, :288, :322, :347, :374
    sig = lambda f: [(p.name, p.default) for p in inspect.signature(f).parameters.values()]  # noqa: E731
    E = inspect.Parameter.empty
    assert sig(R.insert_chunks) == [("user_id", E), ("document_id", E), ("source_filename", E), ("chunks", E), ("org_id", None)]
    assert sig(R.search_knowledge_base) == [("user_id", E), ("query", E), ("limit", 5), ("alpha", 0.5),
                                            ("min_score", 0.0), ("org_id", None)]
    assert sig(R.delete_document_chunks) == [("user_id", E), ("document_id", E)]
    assert sig(R.delete_user_chunks) == [("user_id", E)]
    assert sig(R.get_document_chunk_count) == [("user_id", E), ("document_id", E)]
    assert sig(R.delete_discovery_chunks) == [("org_id", E), ("before", None)]
    assert R.COLLECTION_NAME == "KnowledgeBaseChunk"


def test_insert_search_roundtrip_and_result_shape(kb):
    n = kb.insert_chunks("u1", "doc-a", "runbook.md",
                         _chunks("restart the payment service when latency spikes",
                                 "rotate database credentials every ninety days",
                                 "kafka consumer lag alert runbook"), org_id="org1")
    assert n == 3
    res = kb.search_knowledge_base("u1", "payment service latency", limit=2)
    assert len(res) == 2
    assert set(res[0]) == {"content", "heading_context", "source_filename", "document_id", "chunk_index", "score"}  # :269-276
    assert res[0]["content"].startswith("restart the payment service")
    assert res[0]["source_filename"] == "runbook.md" and res[0]["document_id"] == "doc-a"
    assert res[0]["score"] >= res[1]["score"]
    # alpha=0.5 (default): hybrid ranked fusion, top of both lists -> 0.5/60 + 0.5/60 (:252-259)
    assert res[0]["score"] == pytest.approx(1.0 / 60.0)
    pure = kb.search_knowledge_base("u1", "payment service latency", limit=2, alpha=1.0)
    assert pure[0]["content"] == res[0]["content"] and -1.0 <= pure[0]["score"] <= 1.0 + 1e-6   # cosine


def test_blank_query_and_empty_chunks(kb):
    assert kb.search_knowledge_base("u1", "   ") == []        # :237-238
    assert kb.insert_chunks("u1", "d", "f.md", []) == 0       # :159-160


def test_tenant_scope_is_user_or_org(kb):
    kb.insert_chunks("alice", "d1", "a.md", _chunks("alpha incident postmortem"), org_id="acme")
    kb.insert_chunks("bob", "d2", "b.md", _chunks("alpha incident postmortem"), org_id="acme")
    kb.insert_chunks("carol", "d3", "c.md", _chunks("alpha incident postmortem"), org_id="other")
    own = kb.search_knowledge_base("alice", "alpha incident", limit=10)
    assert {r["document_id"] for r in own} == {"d1"}                              # no org -> user scope only
    shared = kb.search_knowledge_base("alice", "alpha incident", limit=10, org_id="acme")
    assert {r["document_id"] for r in shared} == {"d1", "d2"}                     # :244-249
    assert kb.search_knowledge_base("nobody", "alpha incident", limit=10) == []


def test_min_score_only_filters_when_positive(kb):
    kb.insert_chunks("u", "d", "f.md", _chunks("cpu saturation on api gateway", "completely unrelated gardening tips"))
    all_ = kb.search_knowledge_base("u", "api gateway cpu", limit=5, min_score=0.0)
    assert len(all_) == 2
    some = kb.search_knowledge_base("u", "api gateway cpu", limit=5, alpha=1.0, min_score=0.5)   # :266, cosine scores
    assert [r["content"] for r in some] == ["cpu saturation on api gateway"]
    # hybrid scores are rank-fusion values <= 1/60: the same threshold removes everything, as in the reference
    assert kb.search_knowledge_base("u", "api gateway cpu", limit=5, min_score=0.5) == []
    top = kb.search_knowledge_base("u", "api gateway cpu", limit=5, min_score=0.0166)
    assert [r["content"] for r in top] == ["cpu saturation on api gateway"]


def test_hybrid_keyword_leg_and_alpha(kb):
    """BM25 leg + ranked fusion (weaviate_client.py:252-259): an exact rare keyword wins the sparse list."""
    kb.insert_chunks("u", "d", "f.md", _chunks("generic notes about restarting services and checking dashboards",
                                               "error code zx9981 means the ledger shard is read only",
                                               "more generic notes about services dashboards and restarts"))
    kw = kb.search_knowledge_base("u", "zx9981", limit=3, alpha=0.0)              # keyword only
    assert [r["chunk_index"] for r in kw] == [1] and kw[0]["score"] == pytest.approx(1.0 / 60.0)
    hy = kb.search_knowledge_base("u", "zx9981", limit=3, alpha=0.5)
    assert hy[0]["chunk_index"] == 1 and len(hy) == 3                             # dense leg still contributes the rest
    assert hy[0]["score"] > hy[1]["score"] >= hy[2]["score"] > 0
    # tenant scope applies to the keyword leg too
    kb.insert_chunks("someone-else", "d9", "x.md", _chunks("zx9981 zx9981 zx9981"))
    assert {r["document_id"] for r in kb.search_knowledge_base("u", "zx9981", limit=5, alpha=0.0)} == {"d"}
    # deleting a document removes its postings
    kb.delete_document_chunks("u", "d")
    assert kb.search_knowledge_base("u", "zx9981", limit=5, alpha=0.0) == []


def test_bm25_and_ranked_fusion_units():
    from aurora_b200.bm25 import BM25Index, ranked_fusion, tokenize

    assert tokenize("Kafka-consumer LAG, alert#7!") == ["kafka", "consumer", "lag", "alert", "7"]
    ix = BM25Index()
    ix.add(1, "redis cache eviction policy")
    ix.add(2, "redis redis redis cluster failover")
    ix.add(3, "postgres vacuum tuning")
    top = ix.search("redis failover", 10)
    assert [d for d, _ in top] == [2, 1]                                           # more matches, rarer term
    assert ix.search("redis", 10, allow=lambda d: d != 2)[0][0] == 1
    ix.add(1, "now about kafka")                                                   # upsert replaces postings
    assert [d for d, _ in ix.search("redis", 10)] == [2]
    assert ix.remove(2) and not ix.remove(2) and ix.search("redis", 10) == []
    fused = ranked_fusion([(0.5, [10, 11, 12]), (0.5, [12, 10])], 3)
    assert [d for d, _ in fused] == [10, 12, 11]
    assert fused[0][1] == pytest.approx(0.5 / 60 + 0.5 / 61) and fused[1][1] == pytest.approx(0.5 / 62 + 0.5 / 60)


def test_upsert_is_idempotent_on_user_doc_chunk(kb):
    kb.insert_chunks("u", "d", "f.md", _chunks("first version of the text"))
    kb.insert_chunks("u", "d", "f.md", _chunks("second version of the text"))          # same uuid5 key (:172)
    assert kb.get_document_chunk_count("u", "d") == 1
    res = kb.search_knowledge_base("u", "version of the text", limit=5)
    assert [r["content"] for r in res] == ["second version of the text"]


def test_deletes_and_counts(kb):
    kb.insert_chunks("u", "d1", "f.md", _chunks("one", "two", "three"))
    kb.insert_chunks("u", "d2", "g.md", _chunks("four"))
    assert kb.get_document_chunk_count("u", "d1") == 3
    assert kb.delete_document_chunks("u", "d1") == 3           # :288-319
    assert kb.get_document_chunk_count("u", "d1") == 0
    assert kb.delete_document_chunks("u", "d1") == 0
    assert kb.delete_user_chunks("u") == 1                     # :322-344
    assert kb.search_knowledge_base("u", "four") == []


def test_discovery_chunks_prefix_and_before(kb):
    kb.insert_chunks("u", "discovery:20260101:aa", "topology", _chunks("service a talks to b"), org_id="o")
    kb.insert_chunks("u", "manual-doc", "m.md", _chunks("service a talks to b"), org_id="o")
    assert kb.delete_discovery_chunks("o", before="1999-01-01T00:00:00+00:00") == 0     # created_at < before (:383-384)
    assert kb.delete_discovery_chunks("o") == 1
    assert kb.get_document_chunk_count("u", "manual-doc") == 1


def test_error_conventions_when_the_backend_is_down():
    R.configure(factory=lambda: (_ for _ in ()).throw(RuntimeError("backend down")))
    assert R.search_knowledge_base("u", "q") == []                       # :283-285
    assert R.delete_document_chunks("u", "d") == -1                      # :317-319
    assert R.delete_user_chunks("u") == -1                               # :342-344
    assert R.get_document_chunk_count("u", "d") == 0                     # :369-371
    assert R.delete_discovery_chunks("o") == 0                           # :392-394
    with pytest.raises(RuntimeError):                                    # :210-212 re-raises for Celery retry
        R.insert_chunks("u", "d", "f.md", _chunks("x"))


def test_private_client_facade_used_by_rca_prompt_builder(kb):
    """chat/background/rca_prompt_builder.py:276-317 grabs (client, collection) and runs its own hybrid query."""
    kb.insert_chunks("u", "discovery:20260101:ab", "gke-topology", _chunks("checkout depends on payments and redis"), org_id="o")
    kb.insert_chunks("u", "other", "notes.md", _chunks("checkout depends on payments and redis"), org_id="o")
    client, collection = kb._get_weaviate_client()
    assert client.is_ready()
    f = Filter.by_property("org_id").equal("o") & Filter.by_property("document_id").like("discovery:*")
    resp = collection.query.hybrid(query="checkout payments", limit=3, alpha=0.5, fusion_type=HybridFusion.RANKED,
                                   filters=f, return_metadata=["score"])
    assert len(resp.objects) == 1
    assert resp.objects[0].properties["source_filename"] == "gke-topology"
    assert resp.objects[0].metadata.score > 0


def test_snapshot_roundtrip(tmp_path):
    """save -> load keeps ids, tenant scope, upsert keys, the keyword leg; deleted chunks stay deleted."""
    emb = HashEmbedder(64)
    a = R.KnowledgeBase(emb, capacity=256, index_factory=lambda dim, cap: OracleIndex(dim, cap))
    a.insert("u1", "d1", "a.md", _chunks("redis failover procedure", "postgres vacuum tuning", "kafka lag alert zx77"), "org")
    a.insert("u2", "d2", "b.md", _chunks("unrelated notes"))
    a.delete_where(lambda p: p["document_id"] == "d1" and p["chunk_index"] == 1)
    before = [(o.properties["document_id"], o.properties["chunk_index"], round(o.metadata.score, 6))
              for o in a.query("redis failover", 5, user_id="u1", alpha=0.5)]
    a.save(str(tmp_path / "snap"))
    b = R.KnowledgeBase.load(str(tmp_path / "snap"), emb, capacity=256, index_loader=lambda path, cap: OracleIndex.load(path, cap))
    after = [(o.properties["document_id"], o.properties["chunk_index"], round(o.metadata.score, 6))
             for o in b.query("redis failover", 5, user_id="u1", alpha=0.5)]
    assert after == before and len(after) == 2
    assert [o.properties["chunk_index"] for o in b.query("zx77", 3, user_id="u1", alpha=0.0)] == [2]
    assert b.count_where(lambda p: p["document_id"] == "d1") == 2
    assert b.query("unrelated", 3, user_id="u1", alpha=1.0)[0].properties["document_id"] == "d1"   # u2's chunk stays out of scope
    assert b.insert("u1", "d1", "a.md", _chunks("redis failover procedure v2")) == 1               # same uuid5 key -> upsert
    assert b.count_where(lambda p: p["document_id"] == "d1") == 2


def test_mutation_log_replays_what_a_crash_would_lose(tmp_path):
    """Snapshot, then more inserts / an upsert / deletes logged but never snapshotted, then "crash": a store loaded
    from the old snapshot + the log answers exactly like the store that died; records already inside a snapshot are
    skipped by generation; a torn final line is ignored; save() truncates the log."""
    emb = HashEmbedder(64)
    mk = lambda dim, cap: OracleIndex(dim, cap)   # noqa: E731
    snap, wal = str(tmp_path / "snap"), str(tmp_path / "snap" / "mutations.log")
    os.makedirs(snap)
    a = R.KnowledgeBase(emb, capacity=256, index_factory=mk)
    assert a.attach_wal(wal) == 0
    a.insert("u1", "d1", "a.md", _chunks("redis failover procedure", "postgres vacuum tuning"), "org")
    a.save(snap)
    assert os.path.getsize(wal) == 0                                      # everything logged so far is in the snapshot
    a.insert("u1", "d2", "b.md", _chunks("kafka lag alert zx77", "nginx 502 runbook"))
    a.insert("u1", "d1", "a.md", _chunks("redis failover procedure, second edition"))      # upsert of (u1, d1, 0)
    a.insert("u2", "d9", "z.md", _chunks("another tenant's notes"))
    assert a.delete_where(lambda p: p["document_id"] == "d2" and p["chunk_index"] == 1) == 1
    want = {q: [(o.properties["document_id"], o.properties["chunk_index"], o.properties["content"], round(o.metadata.score, 6))
                for o in a.query(q, 5, user_id="u1", alpha=0.5)] for q in ("redis failover", "zx77 kafka", "nginx")}
    with open(wal, "a") as f:
        f.write('{"op":"put","gen":999,"objs":[["torn')               # the crash hit mid-record
    b = R.KnowledgeBase.load(snap, emb, capacity=256, index_loader=lambda path, cap: OracleIndex.load(path, cap))
    assert b.count_where(lambda p: True) == 2                             # the snapshot alone is stale
    assert b.attach_wal(wal) == 4                                         # three puts + one delete
    got = {q: [(o.properties["document_id"], o.properties["chunk_index"], o.properties["content"], round(o.metadata.score, 6))
               for o in b.query(q, 5, user_id="u1", alpha=0.5)] for q in want}
    assert got == want
    assert b.count_where(lambda p: True) == a.count_where(lambda p: True) == 4 and b.mutations == a.mutations
    assert b.query("another tenant", 3, user_id="u2", alpha=1.0)[0].properties["document_id"] == "d9"
    assert open(wal).read().endswith("\n") and "torn" not in open(wal).read()   # the torn tail is gone: appends stay parseable
    # a second restart from the same (old) snapshot replays again, a restart after a new snapshot replays nothing
    c = R.KnowledgeBase.load(snap, emb, capacity=256, index_loader=lambda path, cap: OracleIndex.load(path, cap))
    assert c.attach_wal(wal) == 4
    c.save(snap)
    d = R.KnowledgeBase.load(snap, emb, capacity=256, index_loader=lambda path, cap: OracleIndex.load(path, cap))
    assert d.attach_wal(wal) == 0 and d.count_where(lambda p: True) == 4


def test_filter_algebra():
    p = {"org_id": "o", "document_id": "discovery:1", "created_at": "2026-01-01T00:00:00+00:00"}
    assert Filter.by_property("org_id").equal("o").matches(p)
    assert not Filter.by_property("org_id").equal("x").matches(p)
    assert Filter.by_property("document_id").like("discovery:*").matches(p)
    assert not Filter.by_property("document_id").like("manual*").matches(p)
    assert Filter.by_property("created_at").less_than("2027").matches(p)
    assert (Filter.by_property("org_id").equal("x") | Filter.by_property("org_id").equal("o")).matches(p)
    assert not (Filter.by_property("org_id").equal("x") & Filter.by_property("org_id").equal("o")).matches(p)


def test_missing_tenant_matches_nothing(kb):
    """The reference always applies ``user_id == u`` (weaviate_client.py:244-249): a missing user and org must
    return [] -- never every tenant's chunks."""
    kb.insert_chunks("alice", "d1", "a.md", _chunks("alpha incident postmortem"), org_id="acme")
    assert kb.search_knowledge_base(None, "alpha incident") == []
    assert kb.search_knowledge_base("", "alpha incident", alpha=1.0) == []
    assert kb.search_knowledge_base("", "alpha incident", alpha=0.0) == []          # keyword leg too
    shared = kb.search_knowledge_base(None, "alpha incident", org_id="acme")          # org scope alone still works
    assert {r["document_id"] for r in shared} == {"d1"}


def test_filters_are_pre_filters_not_post_filters(kb):
    """Weaviate applies `filters` before the vector search.  With > 128 better-matching chunks owned by other
    tenants, a post-filter over the global top-128 would return nothing for the small org."""
    for i in range(20):
        kb.insert_chunks(f"user{i}", f"doc{i}", "big.md",
                         _chunks(*[f"checkout depends on payments and redis replica {j}" for j in range(10)]), org_id="big")
    kb.insert_chunks("u", "discovery:20260101:ab", "gke-topology", _chunks("checkout depends on payments"), org_id="small")
    _, collection = kb._get_weaviate_client()
    f = Filter.by_property("org_id").equal("small") & Filter.by_property("document_id").like("discovery:*")
    for alpha in (1.0, 0.5):
        resp = collection.query.hybrid(query="checkout depends on payments and redis replica", limit=3, alpha=alpha,
                                       fusion_type=HybridFusion.RANKED, filters=f, return_metadata=["score"])
        assert [o.properties["source_filename"] for o in resp.objects] == ["gke-topology"]
    near = collection.query.near_text(query="checkout payments", limit=5, filters=Filter.by_property("org_id").equal("nobody"))
    assert near.objects == []


def test_filter_wire_format_roundtrip():
    f = (Filter.by_property("org_id").equal("o") & Filter.by_property("document_id").like("discovery:*")) | \
        Filter.by_property("created_at").less_than("2027")
    g = Filter.from_json(__import__("json").loads(__import__("json").dumps(f.to_json())))
    for p in ({"org_id": "o", "document_id": "discovery:1"}, {"org_id": "x", "created_at": "2026"}, {"org_id": "x"}):
        assert f.matches(p) == g.matches(p)
    assert (Filter.by_property("org_id").equal("o") & Filter.by_property("document_id").like("d*")).required_equalities() == {"org_id": "o"}
    assert f.required_equalities() == {}            # an OR at the root guarantees nothing
    with pytest.raises(ValueError):
        Filter.from_json(["exec", "x", 1])
