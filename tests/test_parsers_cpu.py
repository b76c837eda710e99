"""Tool-call and reasoning parsers (reference: tests/tool_parsers/*, tests/reasoning/*): full-text and streaming."""
import json

import pytest

from vllm_b200.reasoning import get_reasoning_parser, list_reasoning_parsers
from vllm_b200.tool_parsers import get_tool_parser, list_tool_parsers

ARGS = {"city": "Paris", "days": 3}
CASES = {
    "hermes": 'Let me check.<tool_call>\n{"name": "get_weather", "arguments": {"city": "Paris", "days": 3}}\n</tool_call>',
    "llama3_json": '{"name": "get_weather", "parameters": {"city": "Paris", "days": 3}}',
    "mistral": '[TOOL_CALLS][{"name": "get_weather", "arguments": {"city": "Paris", "days": 3}}]',
    "pythonic": '[get_weather(city="Paris", days=3)]',
    "granite": '<|tool_call|>[{"name": "get_weather", "arguments": {"city": "Paris", "days": 3}}]',
    "jamba": '<tool_calls>[{"name": "get_weather", "arguments": {"city": "Paris", "days": 3}}]</tool_calls>',
    "minimax": '<tool_calls>\n{"name": "get_weather", "arguments": {"city": "Paris", "days": 3}}\n</tool_calls>',
    "phi4_mini_json": 'functools[{"name": "get_weather", "arguments": {"city": "Paris", "days": 3}}]',
    "internlm": 'I will call it.<|action_start|><|plugin|>{"name": "get_weather", "parameters": {"city": "Paris", "days": 3}}<|action_end|>',
    "qwen3_coder": '<tool_call>\n<function=get_weather>\n<parameter=city>\nParis\n</parameter>\n<parameter=days>\n3\n</parameter>\n</function>\n</tool_call>',
    "glm45": '<tool_call>get_weather\n<arg_key>city</arg_key>\n<arg_value>Paris</arg_value>\n<arg_key>days</arg_key>\n<arg_value>3</arg_value>\n</tool_call>',
    "kimi_k2": '<|tool_calls_section_begin|><|tool_call_begin|>functions.get_weather:0<|tool_call_argument_begin|>{"city": "Paris", "days": 3}<|tool_call_end|><|tool_calls_section_end|>',
    "deepseek_v31": '<｜tool▁calls▁begin｜><｜tool▁call▁begin｜>get_weather<｜tool▁sep｜>{"city": "Paris", "days": 3}<｜tool▁call▁end｜><｜tool▁calls▁end｜>',
    "deepseek_v3": '<｜tool▁calls▁begin｜><｜tool▁call▁begin｜>function<｜tool▁sep｜>get_weather\n```json\n{"city": "Paris", "days": 3}\n```<｜tool▁call▁end｜><｜tool▁calls▁end｜>',
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_tool_parser_extract(name):
    p = get_tool_parser(name)()
    res = p.extract(CASES[name])
    assert len(res["tool_calls"]) == 1, res
    fn = res["tool_calls"][0]["function"]
    assert fn["name"] == "get_weather" and json.loads(fn["arguments"]) == ARGS


@pytest.mark.parametrize("name", ["hermes", "jamba", "qwen3_coder", "kimi_k2"])
def test_tool_parser_streaming_emits_once(name):
    p = get_tool_parser(name)()
    text = "thinking... " + CASES[name] if name != "kimi_k2" else CASES[name]
    prev, calls, content = "", [], ""
    for i in range(0, len(text), 7):
        cur = text[: i + 7]
        d = p.extract_streaming(prev, cur, cur[len(prev):])
        prev = cur
        if d:
            calls += d.get("tool_calls", [])
            content += d.get("content", "") or ""
    assert len(calls) == 1 and calls[0]["function"]["name"] == "get_weather"
    assert "<tool_call" not in content and "functions." not in content


def test_registry_sizes():
    assert len(list_tool_parsers()) >= 30 and len(list_reasoning_parsers()) >= 20


@pytest.mark.parametrize("name,text,reasoning,content", [
    ("deepseek_r1", "I think so.</think>\nYes.", "I think so.", "Yes."),
    ("qwen3", "<think>hm</think>\n\nNo.", "hm", "No."),
    ("qwen3", "plain answer", None, "plain answer"),
    ("mistral", "[THINK]plan[/THINK]done", "plan", "done"),
    ("seed_oss", "<seed:think>a b</seed:think>c", "a b", "c"),
    ("hunyuan_a13b", "<think>\nr\n</think>\n<answer>\nfinal\n</answer>", "r", "final"),
    ("gpt_oss", "<|channel|>analysis<|message|>weigh it<|end|><|start|>assistant<|channel|>final<|message|>42<|return|>", "weigh it", "42"),
    ("identity", "<think>x</think>y", None, "<think>x</think>y"),
])
def test_reasoning_extract(name, text, reasoning, content):
    r, c = get_reasoning_parser(name)().extract(text)
    assert r == reasoning and c == content


def test_reasoning_streaming_splits_at_end_tag():
    p = get_reasoning_parser("deepseek_r1")()
    text = "<think>step one, step two</think>The answer."
    prev, r, c = "", "", ""
    for i in range(0, len(text), 5):
        cur = text[: i + 5]
        d = p.extract_streaming(prev, cur, cur[len(prev):])
        prev = cur
        if d:
            r += d.get("reasoning_content", "") or ""
            c += d.get("content", "") or ""
    assert r == "step one, step two" and c == "The answer."


def test_round2_tool_call_formats():
    """MiniMax-M2 / dots invoke-XML, Apertus keyed JSON, Cohere actions, LFM2 marked pythonic, GigaChat-3, GPT-OSS
    harmony commentary calls, Step-3 steptml (reference: vllm/tool_parsers/{minimax_m2,dots,apertus,cohere_command,lfm2,
    gigachat3,gptoss,step3}_tool_parser.py)."""
    import json
    from vllm_b200.tool_parsers import get_tool_parser
    cases = {
        "minimax_m2": ('ok <minimax:tool_call><invoke name="get_weather"><parameter name="city">Paris</parameter>'
                       '<parameter name="days">3</parameter></invoke></minimax:tool_call>', "get_weather",
                       {"city": "Paris", "days": 3}, "ok"),
        "dots": ('<dots_function_call><invoke name=f><parameter name=q>"a b"</parameter></invoke></dots_function_call>',
                 "f", {"q": "a b"}, None),
        "apertus": ('I check. <|tools_prefix|>[{"get_weather": {"city": "Paris"}}]<|tools_suffix|>', "get_weather",
                    {"city": "Paris"}, "I check."),
        "cohere": ('plan <|START_ACTION|>[{"tool_call_id":"0","tool_name":"f","parameters":{"a":1}}]<|END_ACTION|>', "f",
                   {"a": 1}, "plan"),
        "lfm2": ('<|tool_call_start|>[get_weather(location="Paris")]<|tool_call_end|>done', "get_weather",
                 {"location": "Paris"}, "done"),
        "gigachat3": ('thinking<|function_call|>{"name":"f","arguments":{"x":2}}', "f", {"x": 2}, "thinking"),
        "openai": ('<|channel|>commentary to=functions.get_weather <|constrain|>json<|message|>{"city":"SF"}<|call|>',
                   "get_weather", {"city": "SF"}, None),
        "step3p5": ('<｜tool_calls_begin｜><｜tool_call_begin｜>function<｜tool_sep｜><steptml:invoke name="f">'
                    '<steptml:parameter name="a">1</steptml:parameter></steptml:invoke><｜tool_call_end｜><｜tool_calls_end｜>',
                    "f", {"a": 1}, None),
    }
    for name, (text, fn, args, content) in cases.items():
        r = get_tool_parser(name)().extract(text)
        assert len(r["tool_calls"]) == 1, name
        c = r["tool_calls"][0]["function"]
        assert c["name"] == fn and json.loads(c["arguments"]) == args, (name, c)
        assert r["content"] == content, (name, r["content"])
        # plain text passes through untouched, and streaming emits the call exactly once
        assert get_tool_parser(name)().extract("just text")["tool_calls"] == []
        p = get_tool_parser(name)()
        emitted, prev = [], ""
        for i in range(1, len(text) + 1, 7):
            d = p.extract_streaming(prev, text[:i], text[len(prev):i])
            prev = text[:i]
            emitted += (d or {}).get("tool_calls", [])
        d = p.extract_streaming(prev, text, text[len(prev):])
        emitted += (d or {}).get("tool_calls", [])
        assert [e["function"]["name"] for e in emitted] == [fn], (name, emitted)


def test_builtin_chat_templates_and_resolution(tmp_path):
    """Named / per-family chat templates (entrypoints/chat_templates.py): exact wire text of the common dialects, system
    message folding where the dialect has no system role, resolution of --chat-template by name / file / source."""
    from vllm_b200.entrypoints.chat_templates import TEMPLATES, builtin_for_model_type, resolve_chat_template
    from vllm_b200.entrypoints.chat_utils import render_chat

    class Tok:                       # no template of its own, no HF machinery: the jinja path renders
        chat_template = None

    msgs = [{"role": "system", "content": "Be brief."}, {"role": "user", "content": "Hi"},
            {"role": "assistant", "content": "Hello."}, {"role": "user", "content": [{"type": "text", "text": "Bye"}]}]
    r = lambda name, **kw: render_chat(Tok(), msgs, name, **kw)        # noqa: E731
    assert r("chatml") == ("<|im_start|>system\nBe brief.<|im_end|>\n<|im_start|>user\nHi<|im_end|>\n<|im_start|>assistant\n"
                           "Hello.<|im_end|>\n<|im_start|>user\nBye<|im_end|>\n<|im_start|>assistant\n")
    assert r("llama-3").startswith("<|begin_of_text|><|start_header_id|>system<|end_header_id|>\n\nBe brief.<|eot_id|>")
    assert r("llama-3").endswith("<|start_header_id|>user<|end_header_id|>\n\nBye<|eot_id|><|start_header_id|>assistant<|end_header_id|>\n\n")
    assert r("llama-2") == "<s>[INST] <<SYS>>\nBe brief.\n<</SYS>>\n\nHi [/INST] Hello. </s><s>[INST] Bye [/INST]"
    assert r("mistral") == "<s>[INST] Be brief.\n\nHi [/INST]Hello.</s>[INST] Bye [/INST]"
    assert r("gemma") == ("<bos><start_of_turn>user\nBe brief.\n\nHi<end_of_turn>\n<start_of_turn>model\nHello.<end_of_turn>\n"
                          "<start_of_turn>user\nBye<end_of_turn>\n<start_of_turn>model\n")
    assert r("phi-3").endswith("<|user|>\nBye<|end|>\n<|assistant|>\n")
    assert r("deepseek") == "<｜begin▁of▁sentence｜>Be brief.<｜User｜>Hi<｜Assistant｜>Hello.<｜end▁of▁sentence｜><｜User｜>Bye<｜Assistant｜>"
    assert r("command-r").endswith("<|START_OF_TURN_TOKEN|><|USER_TOKEN|>Bye<|END_OF_TURN_TOKEN|><|START_OF_TURN_TOKEN|><|CHATBOT_TOKEN|>")
    assert r("vicuna").startswith("Be brief. USER: Hi ASSISTANT: Hello.</s> USER: Bye ASSISTANT:")
    assert r("alpaca").endswith("### Instruction:\nBye\n\n### Response:\n")
    assert r("chatml", add_generation_prompt=False).endswith("Bye<|im_end|>\n")
    # family fallback when the tokenizer has no template; 32k-vocabulary Llama speaks [INST]
    t = Tok()
    t.model_type = "qwen3"
    assert render_chat(t, msgs[1:2]) == "<|im_start|>user\nHi<|im_end|>\n<|im_start|>assistant\n"
    assert builtin_for_model_type("llama", 128256) is TEMPLATES["llama-3"] and builtin_for_model_type("llama", 32000) is TEMPLATES["llama-2"]
    assert builtin_for_model_type("no_such_family") is None
    # --chat-template: name, file, source, error
    f = tmp_path / "t.jinja"
    f.write_text("{{ messages[0]['content'] }}!")
    assert resolve_chat_template("ChatML") is TEMPLATES["chatml"] and resolve_chat_template(str(f)).endswith("}}!")
    assert resolve_chat_template("{{ 1 }}") == "{{ 1 }}" and resolve_chat_template(None) is None
    with pytest.raises(ValueError):
        resolve_chat_template("nonexistent-template")


def test_round2_reasoning_formats():
    """Tag dialects added in round 2: custom tags, three-token markers with a pre-filled opener, newline-trimmed end tag,
    self-addressed channel framing; batch and streaming agree."""
    def run(name, text, chunks=None):
        p = get_reasoning_parser(name)()
        r, c = p.extract(text)
        ps = get_reasoning_parser(name)()
        rs, cs, cur = "", "", ""
        for d in (chunks or [text[i:i + 5] for i in range(0, len(text), 5)]):
            prev, cur = cur, cur + d
            out = ps.extract_streaming(prev, cur, d)
            if out:
                rs += out.get("reasoning_content") or ""
                cs += out.get("content") or ""
        return (r, c), (rs.strip() or None, cs.strip() or None)

    (r, c), (rs, cs) = run("minimax_m3", "<mm:think>plan the answer</mm:think>The answer is 4.")
    assert (r, c) == ("plan the answer", "The answer is 4.") and (rs, cs) == (r, c)
    (r, c), (rs, cs) = run("kimi_k3", " weigh options <|close|>think<|sep|>Pick B.")           # opener was the generation prefix
    assert (r, c) == ("weigh options", "Pick B.") and (rs, cs) == (r, c)
    (r, c), _ = run("kimi_k3", "<|open|>think<|sep|>x<|close|>think<|sep|>y")
    assert (r, c) == ("x", "y")
    (r, c), (rs, cs) = run("step3p5", "<think>two lines\nof thought\n</think>\nDone.")
    assert (r, c) == ("two lines\nof thought", "Done.") and rs == r and cs == c
    (r, c), _ = run("muse_glimmer", "to=self<|message|>check units<|eom|><|start|>assistant<|message|>12 km<|eot|>")
    assert (r, c) == ("check units", "12 km")
    (r, c), _ = run("muse_glimmer", "<|start|>assistant<|message|>plain reply<|eot|>")
    assert (r, c) == (None, "plain reply")
    for name in ("deepseek_v4", "glm47_moe", "hy_v3", "ling3", "inkling", "poolside_v1", "gptoss"):
        assert name in list_reasoning_parsers()
    assert get_reasoning_parser("hy_v3")().extract("<think>a</think>b") == ("a", "b")


def test_channel_and_function_xml_tool_formats():
    """Kimi-K3 channel markup (typed arguments, escaped attributes, response channel → content) and the
    ``<function name=…><param name=…>`` XML dialect; registry aliases of the remaining reference names."""
    p = get_tool_parser("kimi_k3")()
    text = ('<|open|>response<|sep|>Let me run that.<|close|>response<|sep|><|open|>tools<|sep|>'
            '<|open|>call tool="python" index="1"<|sep|><|open|>argument key="code" type="string"<|sep|>print("a<b")<|close|>argument<|sep|>'
            '<|open|>argument key="opts" type="object"<|sep|>{"n": 3}<|close|>argument<|sep|><|close|>call<|sep|>'
            '<|open|>call tool="say &quot;hi&quot;" index="2"<|sep|><|open|>argument key="times" type="number"<|sep|>2<|close|>argument<|sep|><|close|>call<|sep|>'
            '<|close|>tools<|sep|>')
    out = p.extract(text)
    assert out["content"] == "Let me run that." and len(out["tool_calls"]) == 2
    c0, c1 = out["tool_calls"]
    assert c0["function"]["name"] == "python" and json.loads(c0["function"]["arguments"]) == {"code": 'print("a<b")', "opts": {"n": 3}}
    assert c1["function"]["name"] == 'say "hi"' and json.loads(c1["function"]["arguments"]) == {"times": 2}
    assert p.extract("<|open|>response<|sep|>just text<|close|>response<|sep|>") == {"tool_calls": [], "content": "just text"}
    q = get_tool_parser("minicpm5")()
    out = q.extract('Checking.<function name="get_weather"><param name="city">Paris</param><param name="days">3</param></function>'
                    "<function name='noop'></function>")
    assert out["content"] == "Checking." and [c["function"]["name"] for c in out["tool_calls"]] == ["get_weather", "noop"]
    assert json.loads(out["tool_calls"][0]["function"]["arguments"]) == {"city": "Paris", "days": 3}
    for name in ("cohere_command3", "cohere_command4", "deepseek_v4", "poolside_v1", "inkling", "mimo"):
        assert name in list_tool_parsers()
