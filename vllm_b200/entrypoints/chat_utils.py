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
                elif t in ("image_url", "image", "audio_url", "video_url", "video", "input_audio"):
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
    """Resolution order: explicit template (request / ``--chat-template``: built-in name, file or jinja source) → the
    tokenizer's own template → the built-in template of the model family (``model_type`` kwarg or tokenizer attribute)
    → a neutral role-tag format."""
    from vllm_b200.entrypoints.chat_templates import TEMPLATES, builtin_for_model_type
    msgs, _ = _flatten_content(messages)
    model_type = kwargs.pop("model_type", None) or getattr(tokenizer, "model_type", None)
    if chat_template and chat_template.lower() in TEMPLATES:
        chat_template = TEMPLATES[chat_template.lower()]
    tmpl = chat_template or getattr(tokenizer, "chat_template", None)
    if tmpl is None and not (hasattr(tokenizer, "apply_chat_template") and not hasattr(tokenizer, "backend_tokenizer")):
        tmpl = builtin_for_model_type(model_type, getattr(tokenizer, "vocab_size", None))
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


_ALLOWED_LOCAL_MEDIA_PATH: Any = None    # set by ``--allowed-local-media-path`` / VLLM_ALLOWED_LOCAL_MEDIA_PATH


def set_allowed_local_media_path(path: Any) -> None:
    global _ALLOWED_LOCAL_MEDIA_PATH
    _ALLOWED_LOCAL_MEDIA_PATH = path or None


def resolve_local_media(url: str) -> str:
    """``file://`` URL → real path, only inside the operator-configured directory (reference:
    multimodal/media/connector.py ``_load_file_url``: local files are refused unless
    ``--allowed-local-media-path`` is set, and then confined to that directory after symlink resolution)."""
    import os
    root = _ALLOWED_LOCAL_MEDIA_PATH or os.environ.get("VLLM_ALLOWED_LOCAL_MEDIA_PATH")
    if not root:
        raise ValueError("local file:// media is disabled: start the server with --allowed-local-media-path <dir>")
    root = os.path.realpath(root)
    path = os.path.realpath(url[len("file://"):])
    if os.path.commonpath([root, path]) != root:
        raise ValueError(f"file:// media must live under {root}")
    return path


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
            return Image.open(resolve_local_media(url)).convert("RGB")
        if os.environ.get("VLLM_ALLOW_REMOTE_MEDIA") == "1":
            import requests
            return Image.open(io.BytesIO(requests.get(url, timeout=10).content)).convert("RGB")
        raise ValueError("remote media URLs are disabled (use a data: URL, file://, or set VLLM_ALLOW_REMOTE_MEDIA=1)")
    if t == "video" and part.get("video") is not None:
        return part["video"]
    if t == "video_url":
        url = part["video_url"]["url"] if isinstance(part["video_url"], dict) else part["video_url"]
        return load_video(url, int(part.get("num_frames") or os.environ.get("VLLM_B200_VIDEO_FRAMES", 32)))
    raise ValueError(f"unsupported multimodal content part type {t!r}")


def sample_frame_indices(total: int, num_frames: int) -> list[int]:
    """Uniformly spaced frame indices (reference: multimodal/video.py ``OpenCVVideoBackend`` sampling)."""
    if num_frames <= 0 or total <= num_frames:
        return list(range(total))
    import numpy as np
    return np.linspace(0, total - 1, num_frames).round().astype(int).tolist()


def load_video(url: str, num_frames: int = 32):
    """→ uint8 array [T, H, W, 3].  Accepted: ``data:video/jpeg;base64,<frame>,<frame>,...`` (a list of encoded
    frames), ``data:video/<container>;base64,...`` / ``file://`` containers decoded with OpenCV, ``.npy`` frame arrays."""
    import base64
    import io
    import os
    import tempfile
    import numpy as np
    if url.startswith("data:video/jpeg") or url.startswith("data:video/png"):
        from PIL import Image
        frames = [np.asarray(Image.open(io.BytesIO(base64.b64decode(f))).convert("RGB"))
                  for f in url.split(",")[1:] if f]
        idx = sample_frame_indices(len(frames), num_frames)
        return np.stack([frames[i] for i in idx])
    tmp = None
    if url.startswith("data:"):
        tmp = tempfile.NamedTemporaryFile(suffix="." + url[5:url.index(";")].split("/")[-1], delete=False)
        tmp.write(base64.b64decode(url.split(",", 1)[1]))
        tmp.close()
        path = tmp.name
    elif url.startswith("file://"):
        path = resolve_local_media(url)
    else:
        raise ValueError("remote video URLs are disabled (use a data: URL or file://)")
    try:
        if path.endswith(".npy"):
            arr = np.load(path)
            return arr[sample_frame_indices(arr.shape[0], num_frames)]
        import cv2
        cap = cv2.VideoCapture(path)
        if not cap.isOpened():
            raise ValueError(f"cannot decode video {path!r}")
        total = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        want = set(sample_frame_indices(total, num_frames)) if total > 0 else None
        frames, i = [], 0
        while True:
            ok = cap.grab()
            if not ok:
                break
            if want is None or i in want:
                ok, fr = cap.retrieve()
                if ok:
                    frames.append(cv2.cvtColor(fr, cv2.COLOR_BGR2RGB))
            i += 1
        cap.release()
        if not frames:
            raise ValueError(f"video {path!r} has no decodable frames")
        if want is None:
            frames = [frames[j] for j in sample_frame_indices(len(frames), num_frames)]
        return np.stack(frames)
    finally:
        if tmp is not None:
            os.unlink(tmp.name)


def chat_multi_modal_data(messages: list[dict]) -> Optional[dict]:
    """``{"image": [...], "video": [...]}`` for the media parts of a conversation, in order of appearance (None if
    there are none)."""
    parts = extract_mm_items(messages)
    out: dict[str, list] = {}
    for p in parts:
        t = p.get("type", "")
        if "image" in t:
            out.setdefault("image", []).append(load_mm_item(p))
        elif "video" in t:
            out.setdefault("video", []).append(load_mm_item(p))
    return out or None
