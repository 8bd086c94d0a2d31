"""Drop-in contract of aurora_b200.incident_knowledge vs server/routes/incident_feedback/weaviate_client.py
(Aurora Learn): same names, arguments, return shapes, `similarity = round(1 - distance, 3)`, min_score, org scope,
False / [] / -1 conventions.  CPU only (test doubles for the embedder and the shard)."""

import inspect

import pytest

from aurora_b200 import incident_knowledge as K
from tests.doubles import HashEmbedder, OracleIndex

ORGS = {"alice": "acme", "bob": "acme", "carol": "other"}


@pytest.fixture()
def learn():
    K.configure(encoder=HashEmbedder(96), capacity=512, index_factory=lambda dim, cap: OracleIndex(dim, cap),
                org_resolver=lambda u: ORGS.get(u))
    yield K
    K.configure(factory=lambda: (_ for _ in ()).throw(RuntimeError("backend down")), org_resolver=lambda u: None)


def _store(k, user, incident, title, service="payments", org="acme", summary="connection pool exhausted after deploy"):
    return k.store_good_rca(user, incident, f"fb-{incident}", title, service, "grafana", "critical", summary,
                            [{"content": "checked pool metrics"}, {"content": "rolled back deploy"}],
                            [{"source": "dashboard", "url": "http://x"}], org_id=org)


def test_signatures_match_the_reference_module():
    sig = lambda f: [(p.name, p.default) for p in inspect.signature(f).parameters.values()]  # noqa: E731
    E = inspect.Parameter.empty
    assert sig(K.store_good_rca) == [("user_id", E), ("incident_id", E), ("feedback_id", E), ("alert_title", E),
                                     ("alert_service", E), ("source_type", E), ("severity", E), ("aurora_summary", E),
                                     ("thoughts", E), ("citations", E), ("org_id", None)]                   # :165-177
    assert sig(K.search_similar_good_rcas) == [("user_id", E), ("alert_title", E), ("alert_service", E), ("source_type", E),
                                               ("limit", 2), ("min_score", 0.7)]                           # :246-253
    assert sig(K.delete_incident_knowledge) == [("user_id", E), ("incident_id", E)]
    assert sig(K.delete_user_knowledge) == [("user_id", E)]
    assert K.COLLECTION_NAME == "IncidentKnowledge"


def test_store_and_search_shape_similarity_and_min_score(learn):
    assert _store(learn, "alice", "inc-1", "Payments API p99 latency high")
    assert _store(learn, "bob", "inc-2", "Disk usage above 90 percent on kafka broker", service="kafka",
                  summary="log retention misconfigured")
    hits = learn.search_similar_good_rcas("alice", "Payments API p99 latency high", "payments", "grafana", limit=2, min_score=0.3)
    assert hits and hits[0]["incident_id"] == "inc-1"
    assert set(hits[0]) == {"incident_id", "alert_title", "alert_service", "source_type", "severity", "aurora_summary",
                            "thoughts", "citations", "similarity"}                                         # :309-319
    assert hits[0]["thoughts"] == [{"content": "checked pool metrics"}, {"content": "rolled back deploy"}]   # JSON round trip
    assert hits[0]["citations"][0]["source"] == "dashboard"
    s = hits[0]["similarity"]
    assert s == round(s, 3) and -1.0 <= s <= 1.0                                                           # :318
    assert all(a["similarity"] >= b["similarity"] for a, b in zip(hits, hits[1:]))
    # default min_score = 0.7 drops weak matches (:302)
    assert learn.search_similar_good_rcas("alice", "completely unrelated gardening question", "garden", "manual") == []
    assert len(learn.search_similar_good_rcas("alice", "Payments API p99 latency high", "payments", "grafana", limit=1,
                                              min_score=0.0)) == 1


def test_org_scope_shares_knowledge_and_isolates_other_orgs(learn):
    _store(learn, "alice", "inc-1", "Payments API p99 latency high")
    _store(learn, "carol", "inc-9", "Payments API p99 latency high", org="other")
    bob = learn.search_similar_good_rcas("bob", "Payments API p99 latency high", "payments", "grafana", limit=5, min_score=0.0)
    assert [h["incident_id"] for h in bob] == ["inc-1"]                      # same org as alice (:279-280), not carol's
    carol = learn.search_similar_good_rcas("carol", "Payments API p99 latency high", "payments", "grafana", limit=5, min_score=0.0)
    assert [h["incident_id"] for h in carol] == ["inc-9"]
    # a user without an org falls back to the user_id filter (:281-283)
    learn.store_good_rca("dave", "inc-d", "fb", "Payments API p99 latency high", "payments", "grafana", "low", "s", [], [])
    dave = learn.search_similar_good_rcas("dave", "Payments API p99 latency high", "payments", "grafana", limit=5, min_score=0.0)
    assert [h["incident_id"] for h in dave] == ["inc-d"]


def test_second_rating_replaces_the_first_and_defaults(learn):
    _store(learn, "alice", "inc-1", "Payments API p99 latency high", summary="first summary")
    _store(learn, "alice", "inc-1", "Payments API p99 latency high", summary="second summary")   # uuid5(user:incident), :214
    hits = learn.search_similar_good_rcas("alice", "Payments API p99 latency high", "payments", "grafana", limit=5, min_score=0.0)
    assert [h["aurora_summary"] for h in hits] == ["second summary"]
    learn.store_good_rca("alice", "inc-2", "fb", "Some alert", "", "datadog", "", "s", [], [], org_id="acme")
    h2 = [h for h in learn.search_similar_good_rcas("alice", "Some alert", "", "datadog", limit=5, min_score=0.0) if h["incident_id"] == "inc-2"]
    assert h2[0]["alert_service"] == "unknown" and h2[0]["severity"] == "unknown"                # :222, :224


def test_deletes(learn):
    _store(learn, "alice", "inc-1", "Payments API p99 latency high")
    _store(learn, "alice", "inc-2", "Kafka consumer lag growing", service="kafka")
    _store(learn, "bob", "inc-3", "Kafka consumer lag growing", service="kafka")
    assert learn.delete_incident_knowledge("alice", "inc-1") is True                             # :331-361
    assert learn.delete_incident_knowledge("alice", "inc-1") is True                             # nothing left: still True
    left = learn.search_similar_good_rcas("bob", "Payments API p99 latency high", "payments", "grafana", limit=5, min_score=0.0)
    assert "inc-1" not in [h["incident_id"] for h in left]
    assert learn.delete_user_knowledge("alice") == 1                                             # :364-386
    assert learn.delete_user_knowledge("alice") == 0
    rest = learn.search_similar_good_rcas("bob", "Kafka consumer lag growing", "kafka", "grafana", limit=5, min_score=0.0)
    assert [h["incident_id"] for h in rest] == ["inc-3"]


def test_error_conventions_when_the_backend_is_down():
    K.configure(factory=lambda: (_ for _ in ()).throw(RuntimeError("backend down")), org_resolver=lambda u: None)
    assert K.store_good_rca("u", "i", "f", "t", "s", "src", "sev", "sum", [], []) is False       # :241-243
    assert K.search_similar_good_rcas("u", "t", "s", "src") == []                                # :326-328
    assert K.delete_incident_knowledge("u", "i") is False                                        # :359-361
    assert K.delete_user_knowledge("u") == -1                                                    # :384-386
