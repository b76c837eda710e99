"""Nemotron-H: a stack of single-mixer blocks ``x + mixer(norm(x))`` where the mixer is a Mamba-2 layer, a rope-free GQA
attention layer, a squared-ReLU MLP, or a sparse layer of NON-gated squared-ReLU experts (sigmoid router with selection
bias, group-limited top-k, shared expert, optional latent projection around the routed experts).

Layer kinds come from ``hybrid_override_pattern`` (``M`` Mamba-2, ``*`` attention, ``-`` MLP, ``E`` experts) or
``layers_block_type``.  Reference parity: vllm/model_executor/models/nemotron_h.py, HF ``modeling_nemotron_h``.

The routed experts are plain up → act → down matrices, which the gated grouped kernels of ``layers/moe.py`` do not
cover: they run as a sort-by-expert loop of dense GEMMs (one pair per hit expert).
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn

from vllm_b200 import ops
from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import LogitsProcessor, ParallelLMHead, RMSNorm, VocabParallelEmbedding
from vllm_b200.layers.linear import QKVParallelLinear, ReplicatedLinear, RowParallelLinear
from vllm_b200.layers.mamba import Mamba2Mixer
from vllm_b200.models.mamba2 import _load_mixer_tensor
from vllm_b200.parallel.state import get_tp_group

_KINDS = {"M": "mamba", "*": "attention", "-": "mlp", "E": "moe"}


def _act(name: str):
    if name == "relu2":
        return lambda h: torch.relu(h).square()
    from vllm_b200.layers.common import Activation
    return Activation(name)


class _MLP(nn.Module):
    def __init__(self, H, I, act, bias, dtype, prefix) -> None:
        super().__init__()
        self.up_proj = ReplicatedLinear(H, I, bias, dtype, None, f"{prefix}.up_proj")
        self.down_proj = ReplicatedLinear(I, H, bias, dtype, None, f"{prefix}.down_proj")
        self.act = _act(act)

    def forward(self, x):
        return self.down_proj(self.act(self.up_proj(x)))


class _Attention(nn.Module):
    def __init__(self, cfg, dtype, prefix) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        nkv = cfg.get("num_key_value_heads") or nh
        hd = cfg.get("head_dim") or cfg.get("attention_head_dim") or H // nh
        bias = bool(cfg.get("attention_bias", False))
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nkv, bias, dtype, None, f"{prefix}.qkv_proj")
        self.o_proj = RowParallelLinear(nh * hd, H, bias, dtype=dtype, prefix=f"{prefix}.o_proj")
        self.sizes = [self.qkv_proj.q_size, self.qkv_proj.kv_size, self.qkv_proj.kv_size]
        self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads, prefix=f"{prefix}.attn", rope=None)

    def forward(self, x):
        q, k, v = self.qkv_proj(x).split(self.sizes, dim=-1)
        return self.o_proj(self.attn(q, k, v, None))


class _Experts(nn.Module):
    def __init__(self, cfg, dtype, prefix) -> None:
        super().__init__()
        H, E = cfg["hidden_size"], cfg["n_routed_experts"]
        lat = cfg.get("moe_latent_size")
        Din, I = lat or H, cfg["moe_intermediate_size"]
        bias = bool(cfg.get("mlp_bias", False))
        self.gate = ReplicatedLinear(H, E, False, torch.float32, None, f"{prefix}.gate")
        self.e_score_correction_bias = nn.Parameter(torch.zeros(E, dtype=torch.float32), requires_grad=False)
        self.up = nn.Parameter(torch.empty(E, I, Din, dtype=dtype), requires_grad=False)
        self.down = nn.Parameter(torch.empty(E, Din, I, dtype=dtype), requires_grad=False)
        self.fc1_latent_proj = ReplicatedLinear(H, lat, bias, dtype, None, f"{prefix}.fc1_latent_proj") if lat else None
        self.fc2_latent_proj = ReplicatedLinear(lat, H, bias, dtype, None, f"{prefix}.fc2_latent_proj") if lat else None
        self.shared_experts = _MLP(H, cfg["moe_shared_expert_intermediate_size"], cfg.get("mlp_hidden_act", "relu2"), bias, dtype,
                                   f"{prefix}.shared_experts")
        self.act = _act(cfg.get("mlp_hidden_act", "relu2"))
        self.top_k, self.n_group, self.topk_group = cfg["num_experts_per_tok"], cfg.get("n_group") or 1, cfg.get("topk_group") or 1
        self.renorm, self.scale = bool(cfg.get("norm_topk_prob", True)), float(cfg.get("routed_scaling_factor", 1.0))

    def forward(self, x):
        w, ids = ops.grouped_topk(self.gate(x.float()), self.top_k, self.n_group, self.topk_group, self.renorm, "sigmoid",
                                  self.e_score_correction_bias, self.scale)
        h = self.fc1_latent_proj(x) if self.fc1_latent_proj is not None else x
        T, k = ids.shape
        flat = ids.reshape(-1).long()
        order = flat.argsort(stable=True)
        tok = order // k
        counts = torch.bincount(flat, minlength=self.up.shape[0]).tolist()
        out = torch.zeros(T, h.shape[1], dtype=torch.float32, device=x.device)
        wf = w.reshape(-1).float()
        s = 0
        for e, n in enumerate(counts):
            if n == 0:
                continue
            rows = tok[s:s + n]
            y = self.act(h[rows] @ self.up[e].t()) @ self.down[e].t()
            out.index_add_(0, rows, y.float() * wf[order[s:s + n], None])
            s += n
        out = out.to(x.dtype)
        if self.fc2_latent_proj is not None:
            out = self.fc2_latent_proj(out)
        return out + self.shared_experts(x)


class NemotronHBlock(nn.Module):
    def __init__(self, cfg, dtype, prefix, kind) -> None:
        super().__init__()
        H, eps = cfg["hidden_size"], cfg.get("layer_norm_epsilon", cfg.get("rms_norm_eps", 1e-5))
        self.kind = kind
        self.norm = RMSNorm(H, eps, dtype)
        if kind == "mamba":
            lim = cfg.get("time_step_limit") or (0.0, float("inf"))
            G = cfg.get("n_groups", 8)
            self.mixer = Mamba2Mixer(H, cfg["mamba_num_heads"], cfg["mamba_head_dim"], cfg.get("ssm_state_size", 128),
                                     cfg.get("conv_kernel", 4), G, bool(cfg.get("use_bias", False)), bool(cfg.get("use_conv_bias", True)),
                                     cfg.get("chunk_size", 128), eps, dtype, f"{prefix}.mixer",
                                     time_step_limit=(float(lim[0]), float(lim[1])), norm_groups=G)
        elif kind == "attention":
            self.mixer = _Attention(cfg, dtype, f"{prefix}.mixer")
        elif kind == "mlp":
            self.mixer = _MLP(H, cfg["intermediate_size"], cfg.get("mlp_hidden_act", "relu2"), bool(cfg.get("mlp_bias", False)), dtype,
                              f"{prefix}.mixer")
        elif kind == "moe":
            self.mixer = _Experts(cfg, dtype, f"{prefix}.mixer")
        else:
            raise ValueError(f"unknown Nemotron-H layer kind {kind!r}")

    def forward(self, x):
        return x + self.mixer(self.norm(x))


def layer_kinds(cfg: dict) -> list[str]:
    kinds = cfg.get("layers_block_type") or cfg.get("layer_types")
    if not kinds:
        kinds = [_KINDS[c] for c in cfg["hybrid_override_pattern"]]
    return list(kinds)


class NemotronHForCausalLM(nn.Module):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        assert get_tp_group().world_size == 1, "hybrid Mamba-2 models run replicated (tensor_parallel_size=1)"
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        H = cfg["hidden_size"]
        kinds = layer_kinds(cfg)
        self.embeddings = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.layers = nn.ModuleList([NemotronHBlock(cfg, dtype, f"{prefix}backbone.layers.{i}", k) for i, k in enumerate(kinds)])
        self.norm_f = RMSNorm(H, cfg.get("layer_norm_epsilon", cfg.get("rms_norm_eps", 1e-5)), dtype)
        self.lm_head = ParallelLMHead(cfg["vocab_size"], H, dtype=dtype)
        if cfg.get("tie_word_embeddings", False):
            self.lm_head.tie_weights(self.embeddings)
        self.logits_processor = LogitsProcessor(cfg["vocab_size"])

    def embed_input_ids(self, ids):
        return self.embeddings(ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = inputs_embeds if inputs_embeds is not None else self.embeddings(input_ids)
        for layer in self.layers:
            x = layer(x)
        return self.norm_f(x)

    def compute_logits(self, hidden):
        return self.logits_processor(self.lm_head, hidden)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()
        for name, w in weights:
            n = name
            for pre in ("backbone.", "model."):
                if n.startswith(pre):
                    n = n[len(pre):]
            if n.startswith("mtp."):
                continue
            if n == "lm_head.weight" and self.cfg.get("tie_word_embeddings", False):
                continue
            n = (n.replace("embed_tokens.", "embeddings.").replace(".mixer.gate.weight", ".mixer.gate.weight")
                 .replace(".mixer.gate.e_score_correction_bias", ".mixer.e_score_correction_bias"))
            if n.endswith(".mixer.experts.up_proj") or n.endswith(".mixer.experts.down_proj"):
                t = n.replace(".experts.up_proj", ".up").replace(".experts.down_proj", ".down")
                params[t].data.copy_(w)
                loaded.add(t)
                continue
            if ".mixer.experts." in n:                      # per-expert checkpoint tensors: experts.<e>.{up,down}_proj.weight
                head, tail = n.split(".mixer.experts.", 1)
                e, which = tail.split(".")[0], tail.split(".")[1]
                t = f"{head}.mixer.{'up' if which == 'up_proj' else 'down'}"
                params[t].data[int(e)].copy_(w)
                loaded.add(t)
                continue
            got = _load_mixer_tensor(params, n, w)
            if got:
                loaded.add(got)
                continue
            target, shard = n, None
            for part, sid in (("q_proj", "q"), ("k_proj", "k"), ("v_proj", "v")):
                if f".mixer.{part}." in n:
                    target, shard = n.replace(part, "qkv_proj"), sid
                    break
            if target not in params:
                continue
            mod = modules[target.rsplit(".", 1)[0]]
            loader = getattr(mod, "weight_loader", None)
            loader(params[target], w, shard) if loader is not None and shard is not None else (
                loader(params[target], w) if loader is not None else params[target].data.copy_(w))
            loaded.add(target)
        return loaded
