"""Encoder-decoder seq2seq models: BART (and the same skeleton serves mBART / Whisper-style decoders).

Reference parity: vllm/model_executor/models/bart.py + the V1 encoder-decoder path (cross-attention KV written
once per request — ``CrossAttentionSpec`` in v1/kv_cache_interface.py — encoder run on first schedule).

Here: the encoder runs once when a request is first scheduled (dense bidirectional SDPA over the source tokens),
its per-layer cross-attention K/V are kept per request on the device by the model runner (they never grow, so they
do not go through the paged pool), and the decoder is an ordinary causal stack over the paged self-attention cache
plus one dense cross-attention per layer.
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig
from vllm_b200.forward_context import get_forward_context
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import LogitsProcessor, ParallelLMHead, VocabParallelEmbedding
from vllm_b200.layers.linear import QKVParallelLinear, RowParallelLinear
from vllm_b200.parallel.state import get_tp_group

_ACT = {"gelu": F.gelu, "relu": F.relu, "silu": F.silu, "gelu_new": lambda x: F.gelu(x, approximate="tanh")}


class _EncLayer(nn.Module):
    def __init__(self, H, nh, ffn, act, dtype) -> None:
        super().__init__()
        self.nh = nh
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (nn.Linear(H, H, dtype=dtype) for _ in range(4))
        self.self_attn_layer_norm = nn.LayerNorm(H, dtype=dtype)
        self.fc1, self.fc2 = nn.Linear(H, ffn, dtype=dtype), nn.Linear(ffn, H, dtype=dtype)
        self.final_layer_norm = nn.LayerNorm(H, dtype=dtype)
        self.act = _ACT[act]

    def forward(self, x):                       # [n, H] one sequence, bidirectional
        n, H = x.shape
        q, k, v = (p(x).view(n, self.nh, H // self.nh).transpose(0, 1) for p in (self.q_proj, self.k_proj, self.v_proj))
        a = F.scaled_dot_product_attention(q, k, v).transpose(0, 1).reshape(n, H)
        x = self.self_attn_layer_norm(x + self.out_proj(a))
        return self.final_layer_norm(x + self.fc2(self.act(self.fc1(x))))


class _DecLayer(nn.Module):
    def __init__(self, H, nh, ffn, act, dtype, prefix) -> None:
        super().__init__()
        hd = H // nh
        self.nh, self.hd = nh, hd
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nh, True, dtype, None, f"{prefix}.self_attn.qkv_proj")
        self.out_proj = RowParallelLinear(H, H, True, dtype=dtype, prefix=f"{prefix}.self_attn.out_proj")
        self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads,
                              prefix=f"{prefix}.self_attn.attn")
        self.self_attn_layer_norm = nn.LayerNorm(H, dtype=dtype)
        self.cq, self.ck, self.cv, self.co = (nn.Linear(H, H, dtype=dtype) for _ in range(4))
        self.encoder_attn_layer_norm = nn.LayerNorm(H, dtype=dtype)
        self.fc1, self.fc2 = nn.Linear(H, ffn, dtype=dtype), nn.Linear(ffn, H, dtype=dtype)
        self.final_layer_norm = nn.LayerNorm(H, dtype=dtype)
        self.act = _ACT[act]
        self.layer_idx = 0

    def cross_kv(self, enc: torch.Tensor):
        n, H = enc.shape
        return (self.ck(enc).view(n, self.nh, self.hd).transpose(0, 1).contiguous(),
                self.cv(enc).view(n, self.nh, self.hd).transpose(0, 1).contiguous())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        q, k, v = self.qkv_proj(x).split([self.qkv_proj.q_size] * 3, dim=-1)
        x = self.self_attn_layer_norm(x + self.out_proj(self.attn(q, k, v)))
        ctx = get_forward_context()
        cross = ctx.extra.get("cross")
        if cross is not None:                    # (segment bounds, per-request list of per-layer (K, V))
            cu, kvs = cross
            qh = self.cq(x)
            out = torch.zeros_like(qh)
            for i, kv in enumerate(kvs):
                s, e = cu[i], cu[i + 1]
                if e <= s or kv is None:
                    continue
                K, V = kv[self.layer_idx]
                qq = qh[s:e].view(e - s, self.nh, self.hd).transpose(0, 1)
                out[s:e] = F.scaled_dot_product_attention(qq, K, V).transpose(0, 1).reshape(e - s, -1)
            x = self.encoder_attn_layer_norm(x + self.co(out))
        return self.final_layer_norm(x + self.fc2(self.act(self.fc1(x))))


class BartForConditionalGeneration(nn.Module):
    is_encoder_decoder = True

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        assert get_tp_group().world_size == 1, "the encoder-decoder path runs with tensor_parallel_size=1"
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        H = cfg["d_model"]
        act = cfg.get("activation_function", "gelu")
        self.embed_scale = H ** 0.5 if cfg.get("scale_embedding") else 1.0
        self.shared = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        npos = cfg["max_position_embeddings"] + 2                      # BART's learned positions are offset by 2
        self.enc_pos, self.dec_pos = nn.Embedding(npos, H, dtype=dtype), nn.Embedding(npos, H, dtype=dtype)
        self.enc_ln_emb, self.dec_ln_emb = nn.LayerNorm(H, dtype=dtype), nn.LayerNorm(H, dtype=dtype)
        self.enc_layers = nn.ModuleList([_EncLayer(H, cfg["encoder_attention_heads"], cfg["encoder_ffn_dim"], act, dtype)
                                         for _ in range(cfg["encoder_layers"])])
        self.dec_layers = nn.ModuleList([_DecLayer(H, cfg["decoder_attention_heads"], cfg["decoder_ffn_dim"], act, dtype,
                                                   f"{prefix}model.decoder.layers.{i}") for i in range(cfg["decoder_layers"])])
        for i, l in enumerate(self.dec_layers):
            l.layer_idx = i
        self.lm_head = ParallelLMHead(cfg["vocab_size"], H, dtype=dtype)
        self.lm_head.tie_weights(self.shared)
        self.final_logits_bias = nn.Parameter(torch.zeros(cfg["vocab_size"], dtype=dtype), requires_grad=False)
        self.logits_processor = LogitsProcessor(cfg["vocab_size"])

    # ---- encoder side (called by the model runner once per request)
    @torch.inference_mode()
    def encode(self, ids: torch.Tensor) -> list[tuple[torch.Tensor, torch.Tensor]]:
        pos = torch.arange(ids.numel(), device=ids.device) + 2
        x = self.enc_ln_emb(self.shared(ids) * self.embed_scale + self.enc_pos(pos))
        for l in self.enc_layers:
            x = l(x)
        return [l.cross_kv(x) for l in self.dec_layers]

    # ---- decoder side
    def embed_input_ids(self, ids):
        return self.shared(ids) * self.embed_scale

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = inputs_embeds if inputs_embeds is not None else self.embed_input_ids(input_ids)
        x = self.dec_ln_emb(x + self.dec_pos(positions + 2))
        for l in self.dec_layers:
            x = l(x)
        return x

    def compute_logits(self, hidden):
        return self.logits_processor(self.lm_head, hidden) + self.final_logits_bias.float()

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()
        ren = [("model.encoder.embed_positions.", "enc_pos."), ("model.decoder.embed_positions.", "dec_pos."),
               ("model.encoder.layernorm_embedding.", "enc_ln_emb."), ("model.decoder.layernorm_embedding.", "dec_ln_emb."),
               ("model.encoder.layers.", "enc_layers."), ("model.decoder.layers.", "dec_layers."),
               ("model.shared.", "shared."), (".encoder_attn.q_proj.", ".cq."), (".encoder_attn.k_proj.", ".ck."),
               (".encoder_attn.v_proj.", ".cv."), (".encoder_attn.out_proj.", ".co.")]
        for name, w in weights:
            if name in ("model.encoder.embed_tokens.weight", "model.decoder.embed_tokens.weight", "lm_head.weight"):
                continue                                    # tied to model.shared
            n = name
            for a, b in ren:
                n = n.replace(a, b)
            shard = None
            if n.startswith("dec_layers.") and ".self_attn." in n:
                for part, sid in (("q_proj", "q"), ("k_proj", "k"), ("v_proj", "v")):
                    if f".self_attn.{part}." in n:
                        n, shard = n.replace(f".self_attn.{part}.", ".qkv_proj."), sid
                        break
                else:
                    n = n.replace(".self_attn.out_proj.", ".out_proj.")
            elif n.startswith("enc_layers."):
                n = n.replace(".self_attn.", ".")
            if n not in params:
                continue
            if "." not in n:                                 # top-level parameter (final_logits_bias [1, V])
                params[n].data.copy_(w.reshape(params[n].shape))
                loaded.add(n)
                continue
            mod = modules[n.rsplit(".", 1)[0]]
            loader = getattr(mod, "weight_loader", None)
            loader(params[n], w, shard) if loader is not None else params[n].data.copy_(w)
            loaded.add(n)
        return loaded
