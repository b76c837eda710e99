"""Tool-call parsers: model text → OpenAI ``tool_calls`` (reference: vllm/tool_parsers/*, ≈40 formats).

Implemented families (each covers several reference parser names through aliases):
``hermes`` (<tool_call>{json}</tool_call>: hermes, qwen2.5/3, nous), ``llama3_json`` ({"name":..,"parameters":..}:
llama3.x/4 json), ``mistral`` ([TOOL_CALLS][{...}]), ``pythonic`` ([fn(a=1), ...]: llama4 pythonic, olmo3),
``deepseek_v3`` (<｜tool▁call▁begin｜>...), ``json`` (bare JSON object/array).
Every parser offers ``extract`` (full text) and ``extract_streaming`` (incremental deltas).
"""
from __future__ import annotations

import ast
import json
import re
import uuid
from typing import Any, Optional


def _call(name: str, args: Any, index: int) -> dict:
    if not isinstance(args, str):
        args = json.dumps(args, ensure_ascii=False)
    return {"id": f"chatcmpl-tool-{uuid.uuid4().hex[:24]}", "type": "function", "index": index,
            "function": {"name": name, "arguments": args}}


class ToolParser:
    start_markers: tuple[str, ...] = ()

    def __init__(self, tokenizer=None) -> None:
        self.tokenizer = tokenizer
        self._sent_calls = 0
        self._sent_content = 0

    def extract(self, text: str) -> dict:
        raise NotImplementedError

    def in_tool_call(self, text: str) -> bool:
        return any(m in text for m in self.start_markers)

    def saw_tool_call(self, text: str) -> bool:
        return bool(self.extract(text)["tool_calls"])

    def extract_streaming(self, prev: str, cur: str, delta: str) -> Optional[dict]:
        """Default streaming strategy: content before the first marker streams through; complete
        tool calls are emitted as soon as they parse."""
        cut = min((cur.find(m) for m in self.start_markers if m in cur), default=-1)
        out: dict = {}
        if cut == -1:
            # hold back a possible partial marker at the tail
            safe = len(cur)
            for m in self.start_markers:
                for k in range(1, len(m)):
                    if cur.endswith(m[:k]):
                        safe = min(safe, len(cur) - k)
            if safe > self._sent_content:
                out["content"] = cur[self._sent_content:safe]
                self._sent_content = safe
            return out or None
        if cut > self._sent_content:
            out["content"] = cur[self._sent_content:cut]
            self._sent_content = cut
        calls = self.extract(cur)["tool_calls"]
        complete = calls if self._looks_complete(cur) else calls[:-1] if len(calls) > 1 else []
        if len(complete) > self._sent_calls:
            out["tool_calls"] = complete[self._sent_calls:]
            self._sent_calls = len(complete)
        return out or None

    def _looks_complete(self, text: str) -> bool:
        return True


class HermesToolParser(ToolParser):
    start_markers = ("<tool_call>",)
    _rx = re.compile(r"<tool_call>\s*(.*?)\s*(?:</tool_call>|$)", re.S)

    def extract(self, text: str) -> dict:
        calls = []
        for m in self._rx.finditer(text):
            try:
                obj = json.loads(m.group(1))
                calls.append(_call(obj["name"], obj.get("arguments", obj.get("parameters", {})), len(calls)))
            except Exception:
                continue
        content = text[:text.find("<tool_call>")].strip() if calls else text
        return {"tool_calls": calls, "content": content or None}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith("</tool_call>")


class Llama3JsonToolParser(ToolParser):
    start_markers = ("<|python_tag|>", '{"name"')

    def extract(self, text: str) -> dict:
        body = text.replace("<|python_tag|>", "").strip()
        calls = []
        dec = json.JSONDecoder()
        i = 0
        while i < len(body):
            j = body.find("{", i)
            if j < 0:
                break
            try:
                obj, end = dec.raw_decode(body[j:])
            except ValueError:
                i = j + 1
                continue
            if isinstance(obj, dict) and "name" in obj and ("parameters" in obj or "arguments" in obj):
                calls.append(_call(obj["name"], obj.get("parameters", obj.get("arguments", {})), len(calls)))
            i = j + end
        return {"tool_calls": calls, "content": None if calls else text}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith("}")


class MistralToolParser(ToolParser):
    start_markers = ("[TOOL_CALLS]",)

    def extract(self, text: str) -> dict:
        if "[TOOL_CALLS]" not in text:
            return {"tool_calls": [], "content": text}
        pre, _, body = text.partition("[TOOL_CALLS]")
        calls = []
        try:
            arr = json.loads(body.strip())
            for obj in arr if isinstance(arr, list) else [arr]:
                calls.append(_call(obj["name"], obj.get("arguments", {}), len(calls)))
        except Exception:
            for m in re.finditer(r"(\w+)\s*(\{.*?\})(?=\s*(?:\w+\s*\{|$))", body, re.S):
                try:
                    calls.append(_call(m.group(1), json.loads(m.group(2)), len(calls)))
                except Exception:
                    pass
        return {"tool_calls": calls, "content": pre.strip() or None}

    def _looks_complete(self, text: str) -> bool:
        t = text.rstrip()
        return t.endswith("]") or t.endswith("}")


class PythonicToolParser(ToolParser):
    start_markers = ("[",)

    def extract(self, text: str) -> dict:
        t = text.strip()
        if not (t.startswith("[") and t.endswith("]")):
            return {"tool_calls": [], "content": text}
        try:
            tree = ast.parse(t, mode="eval")
            calls = []
            for node in tree.body.elts:
                if not isinstance(node, ast.Call):
                    return {"tool_calls": [], "content": text}
                name = ast.unparse(node.func)
                args = {kw.arg: ast.literal_eval(kw.value) for kw in node.keywords}
                calls.append(_call(name, args, len(calls)))
            return {"tool_calls": calls, "content": None}
        except Exception:
            return {"tool_calls": [], "content": text}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith("]")


class DeepSeekV3ToolParser(ToolParser):
    start_markers = ("<｜tool▁calls▁begin｜>",)
    _rx = re.compile(r"<｜tool▁call▁begin｜>(?:function)?<｜tool▁sep｜>(?P<name>[^\n]+)\n```json\n(?P<args>.*?)\n```<｜tool▁call▁end｜>",
                     re.S)

    def extract(self, text: str) -> dict:
        calls = [_call(m.group("name").strip(), m.group("args"), i) for i, m in enumerate(self._rx.finditer(text))]
        content = text[:text.find(self.start_markers[0])] if calls else text
        return {"tool_calls": calls, "content": content.strip() or None}

    def _looks_complete(self, text: str) -> bool:
        return "<｜tool▁calls▁end｜>" in text or text.rstrip().endswith("<｜tool▁call▁end｜>")


class JsonToolParser(Llama3JsonToolParser):
    start_markers = ("{", "[")


def _json_calls(obj: Any) -> list[dict]:
    """A JSON value in any of the common shapes → list of tool calls."""
    items = obj if isinstance(obj, list) else [obj]
    out = []
    for it in items:
        if not isinstance(it, dict):
            continue
        fn = it.get("function") if isinstance(it.get("function"), dict) else it
        name = fn.get("name")
        args = fn.get("arguments", fn.get("parameters", {}))
        if name:
            out.append(_call(name, args, len(out)))
    return out


class MarkerJsonToolParser(ToolParser):
    """``<start>[{"name":…,"arguments":…}, …]<end>`` families: granite (``<|tool_call|>`` / ``<tool_call>``), jamba and
    hunyuan (``<tool_calls>…</tool_calls>``), phi4-mini (``functools[…]``), xlam (fenced or bare JSON list),
    internlm (``<|action_start|><|plugin|>{…}<|action_end|>``)."""
    end_markers: tuple[str, ...] = ()
    keep_bracket = False      # the start marker is followed directly by the JSON value

    def extract(self, text: str) -> dict:
        cut = min((text.find(m) for m in self.start_markers if m in text), default=-1)
        if cut < 0:
            return {"tool_calls": [], "content": text}
        m = next(m for m in self.start_markers if text.startswith(m, cut))
        body = text[cut + len(m):]
        for e in self.end_markers:
            if e in body:
                body = body.split(e, 1)[0]
        body = body.strip()
        if body.startswith("```"):
            body = body.strip("`").removeprefix("json").strip()
        calls: list[dict] = []
        try:
            calls = _json_calls(json.loads(body))
        except json.JSONDecodeError:
            # newline-separated objects (minimax) or a truncated list: take every object that parses
            dec = json.JSONDecoder()
            i = 0
            while i < len(body):
                j = body.find("{", i)
                if j < 0:
                    break
                try:
                    obj, n = dec.raw_decode(body[j:])
                except json.JSONDecodeError:
                    break
                calls.extend(_json_calls(obj))
                i = j + n
            for k, c in enumerate(calls):
                c["index"] = k
        return {"tool_calls": calls, "content": text[:cut].strip() or None}

    def _looks_complete(self, text: str) -> bool:
        t = text.rstrip()
        return any(t.endswith(e) for e in self.end_markers) if self.end_markers else t.endswith(("]", "}"))


def _marker_parser(starts: tuple[str, ...], ends: tuple[str, ...] = ()) -> type[MarkerJsonToolParser]:
    return type("MarkerJson_" + re.sub(r"\W", "", starts[0]), (MarkerJsonToolParser,),
                {"start_markers": starts, "end_markers": ends})


class XmlParamToolParser(ToolParser):
    """``<tool_call><function=NAME><parameter=KEY>VALUE</parameter>…</function></tool_call>``
    (qwen3-coder / qwen3_xml / seed_oss / step3 style; values are JSON when they parse, else strings)."""
    start_markers = ("<tool_call>", "<function=")
    _fn = re.compile(r"<function=([^>\s]+)>(.*?)(?:</function>|$)", re.S)
    _par = re.compile(r"<parameter=([^>\s]+)>\s*(.*?)\s*</parameter>", re.S)

    def extract(self, text: str) -> dict:
        calls = []
        for m in self._fn.finditer(text):
            args = {}
            for k, v in self._par.findall(m.group(2)):
                try:
                    args[k] = json.loads(v)
                except json.JSONDecodeError:
                    args[k] = v
            calls.append(_call(m.group(1), args, len(calls)))
        cut = min((text.find(x) for x in self.start_markers if x in text), default=len(text))
        return {"tool_calls": calls, "content": (text[:cut].strip() or None) if calls else text}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith(("</function>", "</tool_call>"))


class Glm4MoeToolParser(ToolParser):
    """``<tool_call>NAME\n<arg_key>k</arg_key>\n<arg_value>v</arg_value>…</tool_call>`` (GLM-4.5 / 4.6)."""
    start_markers = ("<tool_call>",)
    _blk = re.compile(r"<tool_call>\s*([^\n<]+)\s*(.*?)(?:</tool_call>|$)", re.S)
    _kv = re.compile(r"<arg_key>(.*?)</arg_key>\s*<arg_value>(.*?)</arg_value>", re.S)

    def extract(self, text: str) -> dict:
        calls = []
        for m in self._blk.finditer(text):
            args = {}
            for k, v in self._kv.findall(m.group(2)):
                try:
                    args[k.strip()] = json.loads(v)
                except json.JSONDecodeError:
                    args[k.strip()] = v.strip()
            calls.append(_call(m.group(1).strip(), args, len(calls)))
        return {"tool_calls": calls, "content": (text[:text.find("<tool_call>")].strip() or None) if calls else text}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith("</tool_call>")


class KimiK2ToolParser(ToolParser):
    """``<|tool_calls_section_begin|><|tool_call_begin|>functions.NAME:IDX<|tool_call_argument_begin|>{json}<|tool_call_end|>…``"""
    start_markers = ("<|tool_calls_section_begin|>", "<|tool_call_begin|>")
    _rx = re.compile(r"<\|tool_call_begin\|>\s*(?:functions\.)?([\w\-.]+?)(?::\d+)?\s*<\|tool_call_argument_begin\|>\s*(.*?)\s*<\|tool_call_end\|>", re.S)

    def extract(self, text: str) -> dict:
        calls = [_call(m.group(1), m.group(2), i) for i, m in enumerate(self._rx.finditer(text))]
        cut = min((text.find(x) for x in self.start_markers if x in text), default=len(text))
        return {"tool_calls": calls, "content": (text[:cut].strip() or None) if calls else text}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith(("<|tool_call_end|>", "<|tool_calls_section_end|>"))


class DeepSeekV31ToolParser(ToolParser):
    """``<｜tool▁calls▁begin｜><｜tool▁call▁begin｜>NAME<｜tool▁sep｜>{json}<｜tool▁call▁end｜>…``"""
    start_markers = ("<｜tool▁calls▁begin｜>",)
    _rx = re.compile(r"<｜tool▁call▁begin｜>(.*?)<｜tool▁sep｜>(.*?)<｜tool▁call▁end｜>", re.S)

    def extract(self, text: str) -> dict:
        calls = [_call(m.group(1).strip(), m.group(2).strip(), i) for i, m in enumerate(self._rx.finditer(text))]
        return {"tool_calls": calls, "content": (text[:text.find(self.start_markers[0])].strip() or None) if calls else text}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith(("<｜tool▁call▁end｜>", "<｜tool▁calls▁end｜>"))


class InvokeXmlToolParser(ToolParser):
    """``<invoke name="fn"><parameter name="k">v</parameter>…</invoke>`` blocks inside an outer wrapper: MiniMax-M2
    (``<minimax:tool_call>``), dots.llm (``<dots_function_call>``), Claude-style ``<function_calls>``.  Values are JSON
    when they parse, strings otherwise."""
    start_markers = ("<minimax:tool_call>", "<dots_function_call>", "<function_calls>", "<invoke ")
    _inv = re.compile(r"<invoke\s+name\s*=\s*\"?([^\">]+)\"?\s*>(.*?)</invoke>", re.S)
    _par = re.compile(r"<parameter\s+name\s*=\s*\"?([^\">]+)\"?\s*>(.*?)</parameter>", re.S)

    def extract(self, text: str) -> dict:
        calls = []
        for m in self._inv.finditer(text):
            args = {}
            for k, v in self._par.findall(m.group(2)):
                v = v.strip()
                try:
                    args[k.strip()] = json.loads(v)
                except json.JSONDecodeError:
                    args[k.strip()] = v
            calls.append(_call(m.group(1).strip(), args, len(calls)))
        cut = min((text.find(x) for x in self.start_markers if x in text), default=len(text))
        return {"tool_calls": calls, "content": (text[:cut].strip() or None) if calls else text}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith(("</invoke>", "</minimax:tool_call>", "</dots_function_call>", "</function_calls>"))


class KeyedJsonToolParser(ToolParser):
    """``<|tools_prefix|>[{"fn": {args}}, …]<|tools_suffix|>`` (Apertus): each list item maps ONE function name to its
    argument object."""
    start_markers = ("<|tools_prefix|>",)

    def extract(self, text: str) -> dict:
        i = text.find(self.start_markers[0])
        if i < 0:
            return {"tool_calls": [], "content": text}
        body = text[i + len(self.start_markers[0]):].split("<|tools_suffix|>", 1)[0].strip()
        calls = []
        try:
            for item in json.loads(body):
                if isinstance(item, dict):
                    for name, args in item.items():
                        calls.append(_call(name, args, len(calls)))
        except json.JSONDecodeError:
            pass
        return {"tool_calls": calls, "content": text[:i].strip() or None}

    def _looks_complete(self, text: str) -> bool:
        return "<|tools_suffix|>" in text


class CohereActionToolParser(ToolParser):
    """Command-R / Command-A: ``<|START_ACTION|>[{"tool_call_id": "0", "tool_name": "fn", "parameters": {…}}]<|END_ACTION|>``
    (older checkpoints: ``Action: ```json [...]```'' with ``tool_name`` / ``parameters``)."""
    start_markers = ("<|START_ACTION|>", "Action: ```json")

    def extract(self, text: str) -> dict:
        cut = min((text.find(m) for m in self.start_markers if m in text), default=-1)
        if cut < 0:
            return {"tool_calls": [], "content": text}
        m = next(m for m in self.start_markers if text.startswith(m, cut))
        body = text[cut + len(m):]
        for e in ("<|END_ACTION|>", "```"):
            if e in body:
                body = body.split(e, 1)[0]
        calls = []
        try:
            for it in json.loads(body.strip()):
                name = it.get("tool_name") or it.get("name")
                if name:
                    calls.append(_call(name, it.get("parameters", it.get("arguments", {})), len(calls)))
        except (json.JSONDecodeError, AttributeError, TypeError):
            pass
        return {"tool_calls": calls, "content": text[:cut].strip() or None}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith(("<|END_ACTION|>", "```"))


class MarkedPythonicToolParser(PythonicToolParser):
    """Pythonic call lists inside sentinels: LFM2 (``<|tool_call_start|>[fn(a=1)]<|tool_call_end|>``)."""
    start_markers = ("<|tool_call_start|>",)

    def extract(self, text: str) -> dict:
        i = text.find(self.start_markers[0])
        if i < 0:
            return {"tool_calls": [], "content": text}
        body = text[i + len(self.start_markers[0]):].split("<|tool_call_end|>", 1)[0].strip()
        inner = super().extract(body)
        tail = text.split("<|tool_call_end|>", 1)[1].strip() if "<|tool_call_end|>" in text else ""
        content = (text[:i].strip() + (" " + tail if tail else "")).strip() or None
        return {"tool_calls": inner["tool_calls"], "content": content if inner["tool_calls"] else text}

    def _looks_complete(self, text: str) -> bool:
        return "<|tool_call_end|>" in text


class GigaChatToolParser(ToolParser):
    """GigaChat-3: ``function call<|role_sep|>\n{"name": …, "arguments": {…}}`` or ``<|function_call|>{…}`` (one call)."""
    start_markers = ("<|function_call|>", "function call<|role_sep|>")
    _rx = re.compile(r"(?:function call<\|role_sep\|>\n?|<\|function_call\|>)(.*)", re.S)

    def extract(self, text: str) -> dict:
        m = self._rx.search(text)
        if not m:
            return {"tool_calls": [], "content": text}
        body = m.group(1).split("<|message_sep|>", 1)[0].strip()
        try:
            calls = _json_calls(json.loads(body))
        except json.JSONDecodeError:
            calls = []
        return {"tool_calls": calls, "content": (text[:m.start()].strip() or None) if calls else text}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith(("}", "<|message_sep|>"))


class HarmonyToolParser(ToolParser):
    """GPT-OSS harmony format: ``<|start|>assistant<|channel|>commentary to=functions.NAME <|constrain|>json<|message|>{args}<|call|>``;
    text of the ``final`` channel is the content (the ``analysis`` channel is handled by the reasoning parser)."""
    start_markers = ("to=functions.",)
    _rx = re.compile(r"to=functions\.([\w\-.]+)[^<]*(?:<\|constrain\|>\s*\w+\s*)?<\|message\|>(.*?)(?:<\|call\|>|<\|end\|>|$)", re.S)
    _final = re.compile(r"<\|channel\|>final<\|message\|>(.*?)(?:<\|end\|>|<\|return\|>|$)", re.S)

    def extract(self, text: str) -> dict:
        calls = [_call(m.group(1), m.group(2).strip(), i) for i, m in enumerate(self._rx.finditer(text))]
        fm = self._final.search(text)
        content = fm.group(1).strip() if fm else (None if calls else text)
        return {"tool_calls": calls, "content": content or None}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith(("<|call|>", "<|end|>", "<|return|>"))


class Step3ToolParser(ToolParser):
    """Step-3: ``<｜tool_calls_begin｜><｜tool_call_begin｜>function<｜tool_sep｜><steptml:invoke name="fn">
    <steptml:parameter name="k">v</steptml:parameter></steptml:invoke><｜tool_call_end｜><｜tool_calls_end｜>``."""
    start_markers = ("<｜tool_calls_begin｜>", "<steptml:invoke")
    _inv = re.compile(r"<steptml:invoke\s+name=\"([^\"]+)\">(.*?)</steptml:invoke>", re.S)
    _par = re.compile(r"<steptml:parameter\s+name=\"([^\"]+)\">(.*?)</steptml:parameter>", re.S)

    def extract(self, text: str) -> dict:
        calls = []
        for m in self._inv.finditer(text):
            args = {}
            for k, v in self._par.findall(m.group(2)):
                try:
                    args[k] = json.loads(v)
                except json.JSONDecodeError:
                    args[k] = v.strip()
            calls.append(_call(m.group(1), args, len(calls)))
        cut = min((text.find(x) for x in self.start_markers if x in text), default=len(text))
        return {"tool_calls": calls, "content": (text[:cut].strip() or None) if calls else text}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith(("</steptml:invoke>", "<｜tool_call_end｜>", "<｜tool_calls_end｜>"))




class ChannelCallToolParser(ToolParser):
    """Kimi-K3 channel markup: the reply lives in ``<|open|>response<|sep|>…<|close|>response<|sep|>``, tool calls in a
    sibling ``tools`` channel::

        <|open|>tools<|sep|><|open|>call tool="NAME" index="1"<|sep|>
          <|open|>argument key="k" type="string"<|sep|>raw text<|close|>argument<|sep|>
          <|open|>argument key="n" type="number"<|sep|>3<|close|>argument<|sep|>
        <|close|>call<|sep|><|close|>tools<|sep|>

    ``type="string"`` values are taken verbatim, every other type is JSON-decoded; attribute values are entity-escaped
    (``&quot;``, ``&amp;``)."""
    start_markers = ("<|open|>tools<|sep|>", "<|open|>call ")
    _call_re = re.compile(r"<\|open\|>call\s+([^<]*?)<\|sep\|>(.*?)<\|close\|>call<\|sep\|>", re.S)
    _arg_re = re.compile(r"<\|open\|>argument\s+([^<]*?)<\|sep\|>(.*?)<\|close\|>argument<\|sep\|>", re.S)
    _attr_re = re.compile(r"(\w+)\s*=\s*\"([^\"]*)\"")
    _resp_re = re.compile(r"<\|open\|>response<\|sep\|>(.*?)(?:<\|close\|>response<\|sep\|>|$)", re.S)

    @classmethod
    def _attrs(cls, s: str) -> dict:
        return {k: v.replace("&quot;", '"').replace("&amp;", "&") for k, v in cls._attr_re.findall(s)}

    def extract(self, text: str) -> dict:
        calls = []
        for m in self._call_re.finditer(text):
            head = self._attrs(m.group(1))
            args = {}
            for am in self._arg_re.finditer(m.group(2)):
                a = self._attrs(am.group(1))
                raw = am.group(2)
                if a.get("type", "string") == "string":
                    args[a.get("key", "")] = raw
                else:
                    try:
                        args[a.get("key", "")] = json.loads(raw)
                    except json.JSONDecodeError:
                        args[a.get("key", "")] = raw.strip()
            calls.append(_call(head.get("tool", ""), args, len(calls)))
        rm = self._resp_re.search(text)
        if rm:
            content = rm.group(1).strip() or None
        else:
            cut = min((text.find(x) for x in self.start_markers if x in text), default=len(text))
            content = text[:cut].strip() or None
        return {"tool_calls": calls, "content": content if (calls or rm) else text}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith(("<|close|>call<|sep|>", "<|close|>tools<|sep|>"))


class FunctionParamXmlToolParser(ToolParser):
    """``<function name="fn"><param name="k">v</param>…</function>`` blocks (MiniCPM-5 XML dialect); values are JSON when
    they parse, strings otherwise; tokenizer space / newline glyphs (``Ġ`` / ``Ċ``) left in undecoded text are normalised."""
    start_markers = ("<function name=", "<function ")
    _fn = re.compile(r"<function\s+name\s*=\s*['\"]([^'\"]+)['\"][^>]*>(.*?)</function>", re.S)
    _par = re.compile(r"<param\s+name\s*=\s*['\"]([^'\"]+)['\"]\s*>(.*?)</param>", re.S)

    def extract(self, text: str) -> dict:
        text = text.replace("\u0120", " ").replace("\u010a", "\n")
        calls = []
        for m in self._fn.finditer(text):
            args = {}
            for k, v in self._par.findall(m.group(2)):
                v = v.strip()
                try:
                    args[k] = json.loads(v)
                except json.JSONDecodeError:
                    args[k] = v
            calls.append(_call(m.group(1), args, len(calls)))
        cut = min((text.find(x) for x in self.start_markers if x in text), default=len(text))
        return {"tool_calls": calls, "content": (text[:cut].strip() or None) if calls else text}

    def _looks_complete(self, text: str) -> bool:
        return text.rstrip().endswith("</function>")


_PARSERS: dict[str, type[ToolParser]] = {
    "hermes": HermesToolParser, "qwen": HermesToolParser, "qwen3": HermesToolParser, "nous": HermesToolParser,
    "llama3_json": Llama3JsonToolParser, "llama4_json": Llama3JsonToolParser,
    "mistral": MistralToolParser, "pythonic": PythonicToolParser, "llama4_pythonic": PythonicToolParser,
    "olmo3": PythonicToolParser, "deepseek_v3": DeepSeekV3ToolParser,
    "json": JsonToolParser,
    "granite": _marker_parser(("<|tool_call|>", "<tool_call>")), "granite-20b-fc": _marker_parser(("<function_call>",)),
    "jamba": _marker_parser(("<tool_calls>",), ("</tool_calls>",)),
    "hunyuan_a13b": _marker_parser(("<tool_calls>",), ("</tool_calls>",)),
    "minimax": _marker_parser(("<tool_calls>",), ("</tool_calls>",)),
    "phi4_mini_json": _marker_parser(("functools",)),
    "xlam": _marker_parser(("[TOOL_CALLS]", "```json", "[{")),
    "internlm": _marker_parser(("<|action_start|><|plugin|>",), ("<|action_end|>",)),
    "qwen3_coder": XmlParamToolParser, "qwen3_xml": XmlParamToolParser, "seed_oss": XmlParamToolParser,
    "step3": XmlParamToolParser, "glm45": Glm4MoeToolParser, "glm47": Glm4MoeToolParser,
    "kimi_k2": KimiK2ToolParser, "longcat": _marker_parser(("<longcat_tool_call>",), ("</longcat_tool_call>",)),
    "ernie45": _marker_parser(("<tool_call>",), ("</tool_call>",)), "functiongemma": PythonicToolParser, "deepseek_v31": DeepSeekV31ToolParser, "deepseek_v32": DeepSeekV31ToolParser,
    "minimax_m2": InvokeXmlToolParser, "dots": InvokeXmlToolParser, "dots1": InvokeXmlToolParser,
    "apertus": KeyedJsonToolParser, "cohere": CohereActionToolParser, "command_r": CohereActionToolParser,
    "lfm2": MarkedPythonicToolParser, "gigachat3": GigaChatToolParser, "openai": HarmonyToolParser,
    "gptoss": HarmonyToolParser, "step3p5": Step3ToolParser, "step3_xml": Step3ToolParser,
    "hy_v3": _marker_parser(("<tool_calls>",), ("</tool_calls>",)), "ling": HermesToolParser, "ling3": HermesToolParser,
    "kimi_k3": ChannelCallToolParser, "minicpm5": FunctionParamXmlToolParser,
    "cohere_command3": CohereActionToolParser, "cohere_command4": CohereActionToolParser,
    "deepseek_v4": DeepSeekV31ToolParser, "poolside_v1": HermesToolParser, "inkling": HermesToolParser, "mimo": HermesToolParser,
    "granite4": HermesToolParser, "glm4_moe": Glm4MoeToolParser, "minimax_m1": _marker_parser(("<tool_calls>",), ("</tool_calls>",)),
}


def register_tool_parser(name: str, cls: type[ToolParser]) -> None:
    _PARSERS[name] = cls


def get_tool_parser(name: str) -> type[ToolParser]:
    if name not in _PARSERS:
        raise KeyError(f"unknown tool parser {name!r}; available: {sorted(_PARSERS)}")
    return _PARSERS[name]


def list_tool_parsers() -> list[str]:
    return sorted(_PARSERS)
