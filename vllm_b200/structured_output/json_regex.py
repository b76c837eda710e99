"""JSON-schema → regex lowering for the built-in structured-output back-end
(the reference delegates this to outlines/xgrammar; this is a compact subset:
objects with required properties in order, arrays, enums, scalars)."""
from __future__ import annotations

import json
import re
from typing import Any

WS = r"[ ]?"        # compact output by default (the reference back-ends use the same single-optional-space rule)
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
        if "minLength" in schema or "maxLength" in schema:
            lo, hi = int(schema.get("minLength", 0)), schema.get("maxLength")
            ch = STRING[1:-2]                                  # one (possibly escaped) character
            return '"' + ch + "{" + str(lo) + "," + ("" if hi is None else str(int(hi))) + '}"'
        fmt = schema.get("format")
        if fmt == "date":
            return r'"[0-9]{4}-(0[1-9]|1[0-2])-(0[1-9]|[12][0-9]|3[01])"'
        if fmt == "date-time":
            return r'"[0-9]{4}-(0[1-9]|1[0-2])-(0[1-9]|[12][0-9]|3[01])T([01][0-9]|2[0-3]):[0-5][0-9]:[0-5][0-9](\.[0-9]+)?(Z|[+-][0-9]{2}:[0-9]{2})"'
        if fmt == "uuid":
            return r'"[0-9a-f]{8}-[0-9a-f]{4}-[0-9a-f]{4}-[0-9a-f]{4}-[0-9a-f]{12}"'
        return STRING
    if t == "integer":
        if "minimum" in schema and schema["minimum"] >= 0:
            return r"(0|[1-9][0-9]*)"
        return INTEGER
    if t == "number":
        return NUMBER
    if t == "boolean":
        return BOOLEAN
    if t == "null":
        return NULL
    if t == "array":
        item = json_schema_to_regex(schema.get("items", {}), depth + 1)
        lo, hi = int(schema.get("minItems", 0)), schema.get("maxItems")
        if lo == 0 and hi is None:
            return r"\[" + WS + f"(({item})({WS},{WS}({item}))*)?" + WS + r"\]"
        more = "{" + str(max(lo - 1, 0)) + "," + ("" if hi is None else str(max(int(hi) - 1, 0))) + "}"
        body = f"({item})({WS},{WS}({item})){more}"
        return r"\[" + WS + (body if lo > 0 else f"({body})?") + WS + r"\]"
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
