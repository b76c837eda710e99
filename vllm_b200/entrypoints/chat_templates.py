"""Built-in chat templates, by name and by model family.

Reference: vllm/transformers_utils/chat_templates/registry.py (fallback templates for model types whose tokenizer
ships none) and examples/template_*.jinja (named templates users pass to ``--chat-template``).  A checkpoint's own
``tokenizer.chat_template`` always wins; these are used when it has none, or when the operator names one:

    --chat-template chatml | llama-3 | llama-2 | mistral | gemma | phi-3 | zephyr | vicuna | alpaca | deepseek |
                    command-r | falcon | <path to a .jinja file> | <inline jinja>

Every template renders ``messages`` (role / content), honours ``add_generation_prompt`` and tolerates an optional leading
system message; formats whose wire grammar has no system role (Gemma, Mistral v1) fold it into the first user turn.
"""
from __future__ import annotations

import os
from typing import Optional

CHATML = (
    "{% for m in messages %}<|im_start|>{{ m['role'] }}\n{{ m['content'] }}<|im_end|>\n{% endfor %}"
    "{% if add_generation_prompt %}<|im_start|>assistant\n{% endif %}")

LLAMA3 = (
    "{{ '<|begin_of_text|>' }}{% for m in messages %}<|start_header_id|>{{ m['role'] }}<|end_header_id|>\n\n"
    "{{ m['content'] | trim }}<|eot_id|>{% endfor %}"
    "{% if add_generation_prompt %}<|start_header_id|>assistant<|end_header_id|>\n\n{% endif %}")

LLAMA2 = (
    "{% if messages[0]['role'] == 'system' %}{% set sys = messages[0]['content'] %}{% set rest = messages[1:] %}"
    "{% else %}{% set sys = none %}{% set rest = messages %}{% endif %}"
    "{% for m in rest %}{% if m['role'] == 'user' %}<s>[INST] "
    "{% if loop.first and sys %}<<SYS>>\n{{ sys }}\n<</SYS>>\n\n{% endif %}{{ m['content'] | trim }} [/INST]"
    "{% else %} {{ m['content'] | trim }} </s>{% endif %}{% endfor %}")

MISTRAL = (
    "{% if messages[0]['role'] == 'system' %}{% set sys = messages[0]['content'] %}{% set rest = messages[1:] %}"
    "{% else %}{% set sys = none %}{% set rest = messages %}{% endif %}<s>"
    "{% for m in rest %}{% if m['role'] == 'user' %}[INST] {% if loop.first and sys %}{{ sys }}\n\n{% endif %}"
    "{{ m['content'] | trim }} [/INST]{% else %}{{ m['content'] | trim }}</s>{% endif %}{% endfor %}")

GEMMA = (
    "{% if messages[0]['role'] == 'system' %}{% set sys = messages[0]['content'] %}{% set rest = messages[1:] %}"
    "{% else %}{% set sys = none %}{% set rest = messages %}{% endif %}<bos>"
    "{% for m in rest %}<start_of_turn>{{ 'model' if m['role'] == 'assistant' else m['role'] }}\n"
    "{% if loop.first and sys %}{{ sys }}\n\n{% endif %}{{ m['content'] | trim }}<end_of_turn>\n{% endfor %}"
    "{% if add_generation_prompt %}<start_of_turn>model\n{% endif %}")

PHI3 = (
    "{% for m in messages %}<|{{ m['role'] }}|>\n{{ m['content'] }}<|end|>\n{% endfor %}"
    "{% if add_generation_prompt %}<|assistant|>\n{% endif %}")

ZEPHYR = (
    "{% for m in messages %}<|{{ m['role'] }}|>\n{{ m['content'] }}</s>\n{% endfor %}"
    "{% if add_generation_prompt %}<|assistant|>\n{% endif %}")

VICUNA = (
    "{% if messages[0]['role'] == 'system' %}{{ messages[0]['content'] }}{% set rest = messages[1:] %}"
    "{% else %}A chat between a curious user and an artificial intelligence assistant. The assistant gives helpful, "
    "detailed, and polite answers to the user's questions.{% set rest = messages %}{% endif %}"
    "{% for m in rest %}{% if m['role'] == 'user' %} USER: {{ m['content'] | trim }}"
    "{% else %} ASSISTANT: {{ m['content'] | trim }}</s>{% endif %}{% endfor %}"
    "{% if add_generation_prompt %} ASSISTANT:{% endif %}")

ALPACA = (
    "{% for m in messages %}{% if m['role'] == 'system' %}{{ m['content'] }}\n\n"
    "{% elif m['role'] == 'user' %}### Instruction:\n{{ m['content'] | trim }}\n\n"
    "{% else %}### Response:\n{{ m['content'] | trim }}\n\n{% endif %}{% endfor %}"
    "{% if add_generation_prompt %}### Response:\n{% endif %}")

DEEPSEEK = (
    "{{ '<｜begin▁of▁sentence｜>' }}{% for m in messages %}{% if m['role'] == 'system' %}{{ m['content'] }}"
    "{% elif m['role'] == 'user' %}<｜User｜>{{ m['content'] }}"
    "{% else %}<｜Assistant｜>{{ m['content'] }}<｜end▁of▁sentence｜>{% endif %}{% endfor %}"
    "{% if add_generation_prompt %}<｜Assistant｜>{% endif %}")

COMMAND_R = (
    "{{ '<BOS_TOKEN>' }}{% for m in messages %}<|START_OF_TURN_TOKEN|>"
    "{{ '<|SYSTEM_TOKEN|>' if m['role'] == 'system' else ('<|USER_TOKEN|>' if m['role'] == 'user' else '<|CHATBOT_TOKEN|>') }}"
    "{{ m['content'] | trim }}<|END_OF_TURN_TOKEN|>{% endfor %}"
    "{% if add_generation_prompt %}<|START_OF_TURN_TOKEN|><|CHATBOT_TOKEN|>{% endif %}")

FALCON = (
    "{% for m in messages %}{% if m['role'] == 'system' %}{{ m['content'] }}\n"
    "{% elif m['role'] == 'user' %}User: {{ m['content'] | trim }}\n"
    "{% else %}Falcon: {{ m['content'] | trim }}\n{% endif %}{% endfor %}"
    "{% if add_generation_prompt %}Falcon:{% endif %}")

TEMPLATES: dict[str, str] = {
    "chatml": CHATML, "llama-3": LLAMA3, "llama3": LLAMA3, "llama-2": LLAMA2, "llama2": LLAMA2, "mistral": MISTRAL,
    "gemma": GEMMA, "phi-3": PHI3, "phi3": PHI3, "zephyr": ZEPHYR, "vicuna": VICUNA, "alpaca": ALPACA,
    "deepseek": DEEPSEEK, "command-r": COMMAND_R, "cohere": COMMAND_R, "falcon": FALCON,
}

# model_type (config.json) → template used when the tokenizer has none
_BY_MODEL_TYPE: dict[str, str] = {
    "qwen2": "chatml", "qwen2_moe": "chatml", "qwen3": "chatml", "qwen3_moe": "chatml", "qwen3_next": "chatml",
    "qwen2_vl": "chatml", "qwen2_5_vl": "chatml", "internlm2": "chatml", "smollm3": "chatml", "olmo2": "chatml",
    "llama": "llama-3", "llama4": "llama-3", "mistral": "mistral", "mixtral": "mistral",
    "gemma": "gemma", "gemma2": "gemma", "gemma3": "gemma", "gemma3_text": "gemma",
    "phi3": "phi-3", "phimoe": "phi-3", "deepseek_v2": "deepseek", "deepseek_v3": "deepseek",
    "cohere": "command-r", "cohere2": "command-r", "falcon": "falcon",
}


def builtin_for_model_type(model_type: Optional[str], vocab_size: Optional[int] = None) -> Optional[str]:
    name = _BY_MODEL_TYPE.get(model_type or "")
    if name == "llama-3" and vocab_size is not None and vocab_size < 100_000:
        name = "llama-2"                      # 32 k-vocabulary Llama checkpoints speak the [INST] dialect
    return TEMPLATES.get(name) if name else None


def resolve_chat_template(spec: Optional[str]) -> Optional[str]:
    """``--chat-template`` value → jinja source: a built-in name, a file path, or the template text itself."""
    if not spec:
        return None
    if spec.lower() in TEMPLATES:
        return TEMPLATES[spec.lower()]
    if os.path.exists(spec):
        with open(spec) as f:
            return f.read()
    if "{" in spec:
        return spec
    raise ValueError(f"--chat-template {spec!r}: not a built-in template ({', '.join(sorted(set(TEMPLATES)))}), "
                     "not a file, and not jinja source")
