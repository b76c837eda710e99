"""Chat templating (reference: vllm/entrypoints/chat_utils.py, vllm/renderers/)."""
from __future__ import annotations

from typing import Any, Optional

_DEFAULT_TEMPLATE = (
    "{% for m in messages %}<|{{ m['role'] }}|>\n{{ m['content'] }}\n{% endfor %}"
    "{% if add_generation_prompt %}<|assistant|>\n{% endif %}")


def _flatten_content(messages: list[dict]) -> tuple[list[dict], list[Any]]:
    """OpenAI content parts → plain text + extracted multimodal items."""
    out, mm = [], []
    for m in messages:
        c = m.get("content")
        if isinstance(c, list):
            texts = []
            for part in c:
                t = part.get("type")
                if t == "text":
                    texts.append(part.get("text", ""))
                elif t in ("image_url", "image", "audio_url", "video_url", "input_audio"):
                    mm.append(part)
                    texts.append("<image>" if "image" in t else f"<{t.split('_')[0]}>")
            m = {**m, "content": "\n".join(texts)}
        elif c is None:
            m = {**m, "content": ""}
        out.append(m)
    return out, mm


def render_chat(tokenizer, messages: list[dict], chat_template: Optional[str] = None,
                add_generation_prompt: bool = True, continue_final_message: bool = False,
                tools: Optional[list[dict]] = None, **kwargs: Any) -> str:
    msgs, _ = _flatten_content(messages)
    tmpl = chat_template or getattr(tokenizer, "chat_template", None)
    if tmpl is None and hasattr(tokenizer, "apply_chat_template") and not hasattr(tokenizer, "backend_tokenizer"):
        text = tokenizer.apply_chat_template(msgs, add_generation_prompt=add_generation_prompt, tokenize=False)
    elif tmpl is not None and hasattr(tokenizer, "backend_tokenizer"):
        text = tokenizer.apply_chat_template(msgs, chat_template=tmpl, tools=tools, tokenize=False,
                                             add_generation_prompt=add_generation_prompt,
                                             continue_final_message=continue_final_message, **kwargs)
    else:
        import jinja2
        env = jinja2.Environment()
        text = env.from_string(tmpl or _DEFAULT_TEMPLATE).render(
            messages=msgs, add_generation_prompt=add_generation_prompt, tools=tools, **kwargs)
    if continue_final_message and msgs:
        last = msgs[-1]["content"]
        idx = text.rfind(last)
        if idx >= 0:
            text = text[:idx + len(last)]
    return text


def extract_mm_items(messages: list[dict]) -> list[Any]:
    return _flatten_content(messages)[1]
