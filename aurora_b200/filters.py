"""Minimal stand-in for ``weaviate.classes.query.Filter`` / ``HybridFusion``.

The reference builds its tenant and discovery scopes with this algebra
(server/routes/knowledge_base/weaviate_client.py:244-249, :301-304, :380-385;
server/chat/background/rca_prompt_builder.py:286-289): ``by_property(p).equal(v)``,
``.like("prefix*")``, ``.less_than(v)``, combined with ``&`` and ``|``.  Only what those
call sites use is implemented; filters are evaluated against the chunk metadata rows on
the host (string work), never against vectors.  An expression is a plain tree
(``to_json`` / ``from_json``) so it can cross the engine daemon's socket.
"""

from __future__ import annotations

import fnmatch
from typing import Any, Dict, List, Optional, Tuple

_LEAF_OPS = ("eq", "ne", "like", "lt", "gt")


def _eval(tree: List[Any], p: Dict[str, Any]) -> bool:
    op = tree[0]
    if op == "and":
        return _eval(tree[1], p) and _eval(tree[2], p)
    if op == "or":
        return _eval(tree[1], p) or _eval(tree[2], p)
    name, value = tree[1], tree[2]
    got = p.get(name)
    if op == "eq":
        return got == value
    if op == "ne":
        return got != value
    if op == "like":   # Weaviate LIKE: '*' any run of characters, '?' exactly one
        return isinstance(got, str) and fnmatch.fnmatchcase(got, value)
    if op == "lt":
        return got is not None and got < value
    if op == "gt":
        return got is not None and got > value
    raise ValueError(f"unknown filter operator {op!r}")


def _desc(tree: List[Any]) -> str:
    op = tree[0]
    if op in ("and", "or"):
        return f"({_desc(tree[1])} {op.upper()} {_desc(tree[2])})"
    sym = {"eq": "==", "ne": "!=", "like": "LIKE", "lt": "<", "gt": ">"}[op]
    return f"{tree[1]} {sym} {tree[2]!r}"


class _Expr:
    def __init__(self, tree: List[Any]):
        self.tree = tree

    @property
    def desc(self) -> str:
        return _desc(self.tree)

    def matches(self, props: Dict[str, Any]) -> bool:
        return bool(_eval(self.tree, props))

    def __and__(self, other: "_Expr") -> "_Expr":
        return _Expr(["and", self.tree, other.tree])

    def __or__(self, other: "_Expr") -> "_Expr":
        return _Expr(["or", self.tree, other.tree])

    def __repr__(self) -> str:
        return f"Filter[{self.desc}]"

    # ---- wire format (engine daemon) -------------------------------------------------
    def to_json(self) -> List[Any]:
        return self.tree

    @staticmethod
    def from_json(tree: Optional[List[Any]]) -> Optional["_Expr"]:
        if tree is None:
            return None

        def check(t):
            if not isinstance(t, (list, tuple)) or len(t) != 3:
                raise ValueError("malformed filter expression")
            if t[0] in ("and", "or"):
                return [t[0], check(t[1]), check(t[2])]
            if t[0] not in _LEAF_OPS or not isinstance(t[1], str):
                raise ValueError("malformed filter expression")
            return [t[0], t[1], t[2]]

        return _Expr(check(tree))

    # ---- tenant terms ----------------------------------------------------------------
    def required_equalities(self) -> Dict[str, Any]:
        """Property == value terms every match must satisfy (the AND-spine of the tree); used to
        narrow the metadata scan before the full predicate runs."""
        out: Dict[str, Any] = {}

        def walk(t):
            if t[0] == "and":
                walk(t[1]); walk(t[2])
            elif t[0] == "eq":
                out.setdefault(t[1], t[2])

        walk(self.tree)
        return out


class _Property:
    def __init__(self, name: str):
        self.name = name

    def equal(self, value: Any) -> _Expr:
        return _Expr(["eq", self.name, value])

    def not_equal(self, value: Any) -> _Expr:
        return _Expr(["ne", self.name, value])

    def like(self, pattern: str) -> _Expr:
        return _Expr(["like", self.name, pattern])

    def less_than(self, value: Any) -> _Expr:
        return _Expr(["lt", self.name, value])

    def greater_than(self, value: Any) -> _Expr:
        return _Expr(["gt", self.name, value])


class Filter:
    @staticmethod
    def by_property(name: str) -> _Property:
        return _Property(name)

    from_json = staticmethod(_Expr.from_json)


class HybridFusion:
    RANKED = "FUSION_TYPE_RANKED"
    RELATIVE_SCORE = "FUSION_TYPE_RELATIVE_SCORE"
