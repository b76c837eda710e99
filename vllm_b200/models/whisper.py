"""Whisper (speech → text, encoder-decoder).

Reference parity: vllm/model_executor/models/whisper.py (conv front-end + pre-LN encoder, decoder with
cross-attention to the 1500 encoder frames, ``proj_out`` tied to the token embedding) and the V1 encoder-decoder
path (cross-attention K/V computed once per request).

Same execution model as ``models/bart.py``: the model runner calls :meth:`encode` with the request's log-mel
features the first time the request is scheduled and keeps the per-layer cross K/V on the device; the decoder runs
through the paged self-attention cache.
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


class _EncLayer(nn.Module):
    def __init__(self, H, nh, ffn, dtype) -> None:
        super().__init__()
        self.nh = nh
        self.q_proj, self.v_proj, self.out_proj = (nn.Linear(H, H, dtype=dtype) for _ in range(3))
        self.k_proj = nn.Linear(H, H, bias=False, dtype=dtype)
        self.self_attn_layer_norm = nn.LayerNorm(H, dtype=dtype)
        self.fc1, self.fc2 = nn.Linear(H, ffn, dtype=dtype), nn.Linear(ffn, H, dtype=dtype)
        self.final_layer_norm = nn.LayerNorm(H, dtype=dtype)

    def forward(self, x):                                   # [frames, H], bidirectional
        n, H = x.shape
        h = self.self_attn_layer_norm(x)
        q, k, v = (p(h).view(n, self.nh, H // self.nh).transpose(0, 1) for p in (self.q_proj, self.k_proj, self.v_proj))
        x = x + self.out_proj(F.scaled_dot_product_attention(q, k, v).transpose(0, 1).reshape(n, H))
        return x + self.fc2(F.gelu(self.fc1(self.final_layer_norm(x))))


class _DecLayer(nn.Module):
    def __init__(self, H, nh, ffn, dtype, prefix, idx) -> None:
        super().__init__()
        hd = H // nh
        self.nh, self.hd, self.layer_idx = nh, hd, idx
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nh, True, dtype, None, f"{prefix}.self_attn.qkv_proj")
        self.out_proj = RowParallelLinear(H, H, True, dtype=dtype, prefix=f"{prefix}.self_attn.out_proj")
        self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads,
                              prefix=f"{prefix}.self_attn.attn")
        self.self_attn_layer_norm = nn.LayerNorm(H, dtype=dtype)
        self.cq, self.cv, self.co = (nn.Linear(H, H, dtype=dtype) for _ in range(3))
        self.ck = nn.Linear(H, H, bias=False, dtype=dtype)
        self.encoder_attn_layer_norm = nn.LayerNorm(H, dtype=dtype)
        self.fc1, self.fc2 = nn.Linear(H, ffn, dtype=dtype), nn.Linear(ffn, H, dtype=dtype)
        self.final_layer_norm = nn.LayerNorm(H, dtype=dtype)

    def cross_kv(self, enc: torch.Tensor):
        n = enc.shape[0]
        return (self.ck(enc).view(n, self.nh, self.hd).transpose(0, 1).contiguous(),
                self.cv(enc).view(n, self.nh, self.hd).transpose(0, 1).contiguous())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        q, k, v = self.qkv_proj(self.self_attn_layer_norm(x)).split([self.qkv_proj.q_size] * 3, dim=-1)
        x = x + self.out_proj(self.attn(q, k, v))
        cross = get_forward_context().extra.get("cross")
        if cross is not None:
            cu, kvs = cross
            qh = self.cq(self.encoder_attn_layer_norm(x))
            out = torch.zeros_like(qh)
            for i, kv in enumerate(kvs):
                s, e = cu[i], cu[i + 1]
                if e <= s or kv is None:
                    continue
                K, V = kv[self.layer_idx]
                qq = qh[s:e].view(e - s, self.nh, self.hd).transpose(0, 1)
                out[s:e] = F.scaled_dot_product_attention(qq, K, V).transpose(0, 1).reshape(e - s, -1)
            x = x + self.co(out)
        return x + self.fc2(F.gelu(self.fc1(self.final_layer_norm(x))))


class WhisperForConditionalGeneration(nn.Module):
    is_encoder_decoder = True
    encoder_modality = "audio"

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        assert get_tp_group().world_size == 1, "the encoder-decoder path runs with tensor_parallel_size=1"
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        H = cfg["d_model"]
        self.embed_scale = H ** 0.5 if cfg.get("scale_embedding") else 1.0
        self.conv1 = nn.Conv1d(cfg["num_mel_bins"], H, 3, padding=1, dtype=dtype)
        self.conv2 = nn.Conv1d(H, H, 3, stride=2, padding=1, dtype=dtype)
        self.enc_pos = nn.Embedding(cfg["max_source_positions"], H, dtype=dtype)
        self.enc_layers = nn.ModuleList([_EncLayer(H, cfg["encoder_attention_heads"], cfg["encoder_ffn_dim"], dtype)
                                         for _ in range(cfg["encoder_layers"])])
        self.enc_ln = nn.LayerNorm(H, dtype=dtype)
        self.embed_tokens = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.dec_pos = nn.Embedding(cfg["max_target_positions"], H, dtype=dtype)
        self.dec_layers = nn.ModuleList([_DecLayer(H, cfg["decoder_attention_heads"], cfg["decoder_ffn_dim"], dtype,
                                                   f"{prefix}model.decoder.layers.{i}", i)
                                         for i in range(cfg["decoder_layers"])])
        self.dec_ln = nn.LayerNorm(H, dtype=dtype)
        self.lm_head = ParallelLMHead(cfg["vocab_size"], H, dtype=dtype)
        self.lm_head.tie_weights(self.embed_tokens)
        self.logits_processor = LogitsProcessor(cfg["vocab_size"])

    @torch.inference_mode()
    def encode(self, mel: torch.Tensor) -> list[tuple[torch.Tensor, torch.Tensor]]:
        """log-mel [n_mels, frames] → per-decoder-layer cross-attention (K, V) over frames/2 positions."""
        w = self.conv1.weight
        x = F.gelu(self.conv1(mel.to(w.device, w.dtype).unsqueeze(0)))
        x = F.gelu(self.conv2(x))[0].transpose(0, 1)                   # [frames/2, H]
        x = x + self.enc_pos.weight[:x.shape[0]]
        for l in self.enc_layers:
            x = l(x)
        x = self.enc_ln(x)
        return [l.cross_kv(x) for l in self.dec_layers]

    def embed_input_ids(self, ids):
        return self.embed_tokens(ids) * self.embed_scale

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = inputs_embeds if inputs_embeds is not None else self.embed_input_ids(input_ids)
        x = x + self.dec_pos(positions)
        for l in self.dec_layers:
            x = l(x)
        return self.dec_ln(x)

    def compute_logits(self, hidden):
        return self.logits_processor(self.lm_head, hidden)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()
        ren = [("model.encoder.conv", "conv"), ("model.encoder.embed_positions.", "enc_pos."),
               ("model.encoder.layer_norm.", "enc_ln."), ("model.encoder.layers.", "enc_layers."),
               ("model.decoder.embed_tokens.", "embed_tokens."), ("model.decoder.embed_positions.", "dec_pos."),
               ("model.decoder.layer_norm.", "dec_ln."), ("model.decoder.layers.", "dec_layers."),
               (".encoder_attn.q_proj.", ".cq."), (".encoder_attn.k_proj.", ".ck."),
               (".encoder_attn.v_proj.", ".cv."), (".encoder_attn.out_proj.", ".co.")]
        for name, w in weights:
            if name == "proj_out.weight":
                continue                                    # tied to the token embedding
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
            mod = modules[n.rsplit(".", 1)[0]]
            loader = getattr(mod, "weight_loader", None)
            loader(params[n], w, shard) if loader is not None else params[n].data.copy_(w)
            loaded.add(n)
        return loaded
