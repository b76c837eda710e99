"""Shared test helpers: tiny HF models built in-memory (no downloads) and engine factories."""
from __future__ import annotations

import torch

from vllm_b200.config import get_preset_hf_config


def make_hf_model(preset: str, seed: int = 0, dtype=torch.float32):
    """Random-init HuggingFace model with the preset's architecture."""
    import transformers
    cfg = get_preset_hf_config(preset)
    arch = cfg["architectures"][0]
    mt = cfg["model_type"]
    kw = {k: v for k, v in cfg.items() if k not in ("architectures", "model_type", "torch_dtype")}
    hf_cfg = transformers.AutoConfig.for_model(mt, **kw)
    torch.manual_seed(seed)
    model = getattr(transformers, arch)(hf_cfg).to(dtype).eval()
    # router selection biases are zero-initialised buffers: give them values so the biased top-k is actually exercised
    g = torch.Generator().manual_seed(seed + 1)
    for name, buf in model.named_buffers():
        if name.endswith("e_score_correction_bias") or name.endswith(".expert_bias"):
            buf.copy_((torch.rand(buf.shape, generator=g) - 0.5) * 0.2)
    for name, prm in model.named_parameters():
        if name.endswith(".expert_bias"):
            prm.data.copy_((torch.rand(prm.shape, generator=g) - 0.5) * 0.2)
    return model


def make_llm(preset: str, weights=None, **kw):
    from vllm_b200 import LLM
    defaults = dict(device="cpu", max_model_len=256, max_num_seqs=16, max_num_batched_tokens=128,
                    kv_cache_memory_bytes=16 << 20, dtype="float32")
    defaults.update(kw)
    llm = LLM(preset, **defaults)
    if weights is not None:
        llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model.load_weights(weights)
    return llm


def hf_greedy(model, prompt_ids: list[int], n: int) -> list[int]:
    ids = torch.tensor([prompt_ids])
    with torch.no_grad():
        out = model.generate(ids, max_new_tokens=n, do_sample=False, pad_token_id=0,
                             eos_token_id=None)
    return out[0, len(prompt_ids):].tolist()


def llama_full_weights(cfg: dict, seed: int = 0, scale: float = 0.05):
    """Deterministic *unsharded* Llama checkpoint (HF names) so that different TP sizes load the same model."""
    H, I, V, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"], cfg["num_hidden_layers"]
    nh, nkv = cfg["num_attention_heads"], cfg.get("num_key_value_heads") or cfg["num_attention_heads"]
    D = cfg.get("head_dim") or H // nh
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape):
        return (torch.rand(*shape, generator=g) * 2 - 1) * scale

    def norm():
        return 1.0 + (torch.rand(H, generator=g) * 2 - 1) * 0.05

    yield "model.embed_tokens.weight", rnd(V, H)
    for i in range(L):
        p = f"model.layers.{i}."
        yield p + "input_layernorm.weight", norm()
        yield p + "self_attn.q_proj.weight", rnd(nh * D, H)
        yield p + "self_attn.k_proj.weight", rnd(nkv * D, H)
        yield p + "self_attn.v_proj.weight", rnd(nkv * D, H)
        yield p + "self_attn.o_proj.weight", rnd(H, nh * D)
        yield p + "post_attention_layernorm.weight", norm()
        yield p + "mlp.gate_proj.weight", rnd(I, H)
        yield p + "mlp.up_proj.weight", rnd(I, H)
        yield p + "mlp.down_proj.weight", rnd(H, I)
    yield "model.norm.weight", norm()
    yield "lm_head.weight", rnd(V, H)
