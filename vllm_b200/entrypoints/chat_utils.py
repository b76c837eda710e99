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


def load_mm_item(part: dict) -> Any:
    """One OpenAI content part → a decoded media object (images: PIL.Image; tensors / arrays pass through).
    ``data:`` URLs and local ``file://`` paths are supported; remote fetches need network access and are refused
    unless ``VLLM_ALLOW_REMOTE_MEDIA=1``."""
    import base64
    import io
    import os
    t = part.get("type")
    if t == "image" and part.get("image") is not None:
        return part["image"]
    if t == "image_url":
        url = part["image_url"]["url"] if isinstance(part["image_url"], dict) else part["image_url"]
        from PIL import Image
        if url.startswith("data:"):
            return Image.open(io.BytesIO(base64.b64decode(url.split(",", 1)[1]))).convert("RGB")
        if url.startswith("file://"):
            return Image.open(url[len("file://"):]).convert("RGB")
        if os.environ.get("VLLM_ALLOW_REMOTE_MEDIA") == "1":
            import requests
            return Image.open(io.BytesIO(requests.get(url, timeout=10).content)).convert("RGB")
        raise ValueError("remote media URLs are disabled (use a data: URL, file://, or set VLLM_ALLOW_REMOTE_MEDIA=1)")
    raise ValueError(f"unsupported multimodal content part type {t!r}")


def chat_multi_modal_data(messages: list[dict]) -> Optional[dict]:
    """``{"image": [...]}`` for the image parts of a conversation, in order of appearance (None if there are none)."""
    items = [load_mm_item(p) for p in extract_mm_items(messages) if "image" in p.get("type", "")]
    return {"image": items} if items else None
