"""Minimal stand-in for ``weaviate.classes.query.Filter`` / ``HybridFusion``.

The reference builds its tenant and discovery scopes with this algebra
(server/routes/knowledge_base/weaviate_client.py:244-249, :301-304, :380-385;
server/chat/background/rca_prompt_builder.py:286-289): ``by_property(p).equal(v)``,
``.like("prefix*")``, ``.less_than(v)``, combined with ``&`` and ``|``.  Only what those
call sites use is implemented; filters are evaluated against the chunk metadata rows on
the host (string work), never against vectors.
"""

from __future__ import annotations

import fnmatch
from typing import Any, Callable, Dict


class _Expr:
    def __init__(self, fn: Callable[[Dict[str, Any]], bool], desc: str):
        self._fn, self.desc = fn, desc

    def matches(self, props: Dict[str, Any]) -> bool:
        return bool(self._fn(props))

    def __and__(self, other: "_Expr") -> "_Expr":
        return _Expr(lambda p: self.matches(p) and other.matches(p), f"({self.desc} AND {other.desc})")

    def __or__(self, other: "_Expr") -> "_Expr":
        return _Expr(lambda p: self.matches(p) or other.matches(p), f"({self.desc} OR {other.desc})")

    def __repr__(self) -> str:
        return f"Filter[{self.desc}]"


class _Property:
    def __init__(self, name: str):
        self.name = name

    def equal(self, value: Any) -> _Expr:
        return _Expr(lambda p: p.get(self.name) == value, f"{self.name} == {value!r}")

    def not_equal(self, value: Any) -> _Expr:
        return _Expr(lambda p: p.get(self.name) != value, f"{self.name} != {value!r}")

    def like(self, pattern: str) -> _Expr:
        # Weaviate LIKE: '*' any run of characters, '?' exactly one
        return _Expr(lambda p: isinstance(p.get(self.name), str) and fnmatch.fnmatchcase(p[self.name], pattern),
                     f"{self.name} LIKE {pattern!r}")

    def less_than(self, value: Any) -> _Expr:
        return _Expr(lambda p: p.get(self.name) is not None and p[self.name] < value, f"{self.name} < {value!r}")

    def greater_than(self, value: Any) -> _Expr:
        return _Expr(lambda p: p.get(self.name) is not None and p[self.name] > value, f"{self.name} > {value!r}")


class Filter:
    @staticmethod
    def by_property(name: str) -> _Property:
        return _Property(name)


class HybridFusion:
    RANKED = "FUSION_TYPE_RANKED"
    RELATIVE_SCORE = "FUSION_TYPE_RELATIVE_SCORE"
