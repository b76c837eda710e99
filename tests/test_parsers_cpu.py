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
