"""JSON-schema → regex lowering for the built-in structured-output back-end
(the reference delegates this to outlines/xgrammar; this is a compact subset:
objects with required properties in order, arrays, enums, scalars)."""
from __future__ import annotations

import json
import re
from typing import Any

WS = r"[ \n\t]*"
STRING = r'"([^"\\\x00-\x1f]|\\["\\/bfnrt]|\\u[0-9a-fA-F]{4})*"'
INTEGER = r"-?(0|[1-9][0-9]*)"
NUMBER = INTEGER + r"(\.[0-9]+)?([eE][+-]?[0-9]+)?"
BOOLEAN = r"(true|false)"
NULL = r"null"


def json_schema_to_regex(schema: dict[str, Any], depth: int = 0) -> str:
    if depth > 6:
        return _any_value(0)
    if "enum" in schema:
        return "(" + "|".join(re.escape(json.dumps(v)) for v in schema["enum"]) + ")"
    if "const" in schema:
        return re.escape(json.dumps(schema["const"]))
    if "anyOf" in schema or "oneOf" in schema:
        alts = schema.get("anyOf") or schema.get("oneOf")
        return "(" + "|".join(json_schema_to_regex(a, depth + 1) for a in alts) + ")"
    t = schema.get("type")
    if isinstance(t, list):
        return "(" + "|".join(json_schema_to_regex({**schema, "type": x}, depth + 1) for x in t) + ")"
    if t == "string":
        if "pattern" in schema:
            return '"' + schema["pattern"].strip("^$") + '"'
        return STRING
    if t == "integer":
        return INTEGER
    if t == "number":
        return NUMBER
    if t == "boolean":
        return BOOLEAN
    if t == "null":
        return NULL
    if t == "array":
        item = json_schema_to_regex(schema.get("items", {}), depth + 1)
        return r"\[" + WS + f"(({item})({WS},{WS}({item}))*)?" + WS + r"\]"
    if t == "object" or "properties" in schema:
        props = schema.get("properties")
        if not props:
            return r"\{" + WS + f"({STRING}{WS}:{WS}{_any_value(depth + 1)}({WS},{WS}{STRING}{WS}:{WS}{_any_value(depth + 1)})*)?" + WS + r"\}"
        required = set(schema.get("required", list(props)))
        parts = []
        first = True
        for name, sub in props.items():
            kv = re.escape(json.dumps(name)) + WS + ":" + WS + json_schema_to_regex(sub, depth + 1)
            piece = (("" if first else WS + "," + WS) + kv)
            if name in required:
                parts.append(piece)
                first = False
            else:
                parts.append(f"({piece})?") if not first else parts.append(f"({kv}{WS},{WS})?")
        return r"\{" + WS + "".join(parts) + WS + r"\}"
    return _any_value(depth + 1)


def _any_value(depth: int) -> str:
    scalars = f"({STRING}|{NUMBER}|{BOOLEAN}|{NULL})"
    if depth > 2:
        return scalars
    inner = _any_value(depth + 1)
    arr = r"\[" + WS + f"({inner}({WS},{WS}{inner})*)?" + WS + r"\]"
    obj = r"\{" + WS + f"({STRING}{WS}:{WS}{inner}({WS},{WS}{STRING}{WS}:{WS}{inner})*)?" + WS + r"\}"
    return f"({STRING}|{NUMBER}|{BOOLEAN}|{NULL}|{arr}|{obj})"
