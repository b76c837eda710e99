"""Speech-in, text-out decoders: a Whisper-style audio tower whose outputs are projected into the text embedding space and
spliced over ``<audio>`` placeholder positions of an ordinary decoder (no cross-attention).

  Qwen2-Audio — tower + stride-2 average pool + LayerNorm, one linear projector, Qwen2 decoder; frames beyond the clip's
                real length are masked inside the tower and produce no tokens              (reference: models/qwen2_audio.py)
  Voxtral     — tower (always the full 30 s window), ``intermediate_size / hidden`` consecutive frames concatenated per
                token, 2-layer GELU projector, Llama / Mistral decoder                     (reference: models/voxtral.py)

The tower is the encoder half of ``models/whisper.py`` (same layer class); it is replicated per TP rank.  The front-end
(``AudioLMProcessor``) turns a waveform / file / precomputed log-mel into the padded window + valid frame count and
expands each placeholder to the number of tokens the tower will emit.
"""
from __future__ import annotations

from typing import Any, Iterable

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig


class _Layer(nn.Module):
    def __init__(self, H, nh, ffn, dtype) -> None:
        super().__init__()
        self.nh = nh
        self.q_proj, self.v_proj, self.out_proj = (nn.Linear(H, H, dtype=dtype) for _ in range(3))
        self.k_proj = nn.Linear(H, H, bias=False, dtype=dtype)
        self.self_attn_layer_norm = nn.LayerNorm(H, dtype=dtype)
        self.fc1, self.fc2 = nn.Linear(H, ffn, dtype=dtype), nn.Linear(ffn, H, dtype=dtype)
        self.final_layer_norm = nn.LayerNorm(H, dtype=dtype)

    def forward(self, x, mask=None):                        # [frames, H]; mask [1, 1, frames] additive over keys
        n, H = x.shape
        h = self.self_attn_layer_norm(x)
        q, k, v = (p(h).view(n, self.nh, H // self.nh).transpose(0, 1) for p in (self.q_proj, self.k_proj, self.v_proj))
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
        x = x + self.out_proj(a.transpose(0, 1).reshape(n, H))
        return x + self.fc2(F.gelu(self.fc1(self.final_layer_norm(x))))


class AudioTower(nn.Module):
    def __init__(self, ac: dict, dtype, pool: bool) -> None:
        super().__init__()
        H = ac.get("d_model") or ac["hidden_size"]
        nh = ac.get("encoder_attention_heads") or ac["num_attention_heads"]
        ffn = ac.get("encoder_ffn_dim") or ac["intermediate_size"]
        n = ac.get("encoder_layers") or ac["num_hidden_layers"]
        self.conv1 = nn.Conv1d(ac["num_mel_bins"], H, 3, padding=1, dtype=dtype)
        self.conv2 = nn.Conv1d(H, H, 3, stride=2, padding=1, dtype=dtype)
        self.embed_positions = nn.Embedding(ac["max_source_positions"], H, dtype=dtype)
        self.layers = nn.ModuleList([_Layer(H, nh, ffn, dtype) for _ in range(n)])
        self.layer_norm = nn.LayerNorm(H, dtype=dtype)
        self.pool = pool

    def forward(self, mel: torch.Tensor, valid_frames: int | None = None) -> torch.Tensor:
        """mel [n_mels, T] (T = 2 · max_source_positions) → [T/2 (or T/4 pooled), H]; keys past the valid length are masked."""
        w = self.conv1.weight
        x = F.gelu(self.conv1(mel.to(w.device, w.dtype).unsqueeze(0)))
        x = F.gelu(self.conv2(x))[0].transpose(0, 1)
        x = x + self.embed_positions.weight[: x.shape[0]]
        mask = None
        if valid_frames is not None and (valid_frames - 1) // 2 + 1 < x.shape[0]:
            keep = (valid_frames - 1) // 2 + 1
            mask = torch.zeros(1, 1, x.shape[0], dtype=x.dtype, device=x.device)
            mask[..., keep:] = float("-inf")
        for layer in self.layers:
            x = layer(x, mask)
        if self.pool:
            x = F.avg_pool1d(x.t().unsqueeze(0), 2, 2)[0].t()
        return self.layer_norm(x)


class _AudioLM(nn.Module):
    supports_multimodal = True
    pool = False
    default_text = "LlamaForCausalLM"

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        from vllm_b200.models.registry import text_decoder_cls
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        self.audio_token = cfg.get("audio_token_index", cfg.get("audio_token_id"))
        self.audio_tower = AudioTower(cfg["audio_config"], dtype, self.pool)
        self._build_projector(cfg, dtype)
        if cfg.get("text_model_type") == "qwen2":
            cfg.setdefault("attention_bias", True)
        self.language_model = text_decoder_cls(cfg, self.default_text)(vllm_config, prefix=f"{prefix}language_model.")

    @torch.inference_mode()
    def embed_multimodal_item(self, f: dict) -> torch.Tensor:
        mel = torch.as_tensor(f["data"])
        return self._project(self.audio_tower(mel, f.get("valid_frames")))[: f["length"]]

    def embed_input_ids(self, input_ids):
        return self.language_model.embed_input_ids(input_ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        return self.language_model(input_ids, positions, intermediate_tensors, inputs_embeds)

    def compute_logits(self, hidden):
        return self.language_model.compute_logits(hidden)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        own = dict(self.named_parameters())
        text, loaded = [], set()
        for name, w in weights:
            n = name[len("model."):] if name.startswith("model.") else name
            if n.startswith("language_model."):
                rest = n[len("language_model."):]
                rest = rest[len("model."):] if rest.startswith("model.") else rest
                text.append((rest if rest.startswith("lm_head") else "model." + rest, w))
            elif n.startswith("lm_head."):
                text.append((n, w))
            else:
                t = n.replace(".self_attn.", ".")
                if t in own:
                    own[t].data.copy_(w)
                    loaded.add(t)
        loaded |= {"language_model." + x for x in self.language_model.load_weights(text)}
        return loaded


class Qwen2AudioForConditionalGeneration(_AudioLM):
    pool = True
    default_text = "Qwen2ForCausalLM"

    def _build_projector(self, cfg, dtype) -> None:
        self.multi_modal_projector = nn.ModuleDict({"linear": nn.Linear(cfg["audio_config"]["d_model"], cfg["hidden_size"], dtype=dtype)})

    def _project(self, x):
        return self.multi_modal_projector["linear"](x)

    @staticmethod
    def num_tokens(ac: dict, valid_frames: int) -> int:
        after_conv = (valid_frames - 1) // 2 + 1
        return (after_conv - 2) // 2 + 1


class VoxtralForConditionalGeneration(_AudioLM):
    def _build_projector(self, cfg, dtype) -> None:
        ac = cfg["audio_config"]
        self.group = ac["intermediate_size"]
        self.multi_modal_projector = nn.ModuleDict({"linear_1": nn.Linear(self.group, cfg["hidden_size"], bias=False, dtype=dtype),
                                                    "linear_2": nn.Linear(cfg["hidden_size"], cfg["hidden_size"], bias=False, dtype=dtype)})

    def _project(self, x):
        p = self.multi_modal_projector
        return p["linear_2"](F.gelu(p["linear_1"](x.reshape(-1, self.group))))

    @torch.inference_mode()
    def embed_multimodal_item(self, f: dict) -> torch.Tensor:
        mel = torch.as_tensor(f["data"])                     # one or more full windows side by side: [n_mels, k · T]
        T = 2 * self.cfg["audio_config"]["max_source_positions"]
        outs = [self._project(self.audio_tower(mel[:, i:i + T])) for i in range(0, mel.shape[1], T)]
        return torch.cat(outs, dim=0)[: f["length"]]

    @staticmethod
    def num_tokens(ac: dict, valid_frames: int) -> int:
        T = 2 * ac["max_source_positions"]
        windows = max(1, -(-valid_frames // T))
        return windows * ac["max_source_positions"] * (ac.get("hidden_size") or ac["d_model"]) // ac["intermediate_size"]


class UltravoxModel(_AudioLM):
    """Ultravox: Whisper encoder (padding masked) → 8 consecutive frames stacked per token → RMSNorm → linear → SwiGLU →
    (RMSNorm) → linear → (RMSNorm) → Llama-family decoder; v0.5+ checkpoints norm in the middle, earlier ones at the end
    (reference: models/ultravox.py ``UltravoxFeedForwardProjector``)."""

    def _build_projector(self, cfg, dtype) -> None:
        from vllm_b200.layers.common import RMSNorm
        ac = cfg["audio_config"]
        self.stack = int(cfg.get("stack_factor", 8))
        d_in = (ac.get("d_model") or ac["hidden_size"]) * self.stack
        mid = int(cfg.get("projector_hidden_size", cfg["hidden_size"]))
        self.swiglu = cfg.get("projector_act", "swiglu") == "swiglu"
        d_mid = mid // 2 if self.swiglu else mid
        ln_mid = bool(cfg.get("projector_ln_mid", False))
        mods = {"ln_pre": RMSNorm(d_in, 1e-6, dtype), "linear_1": nn.Linear(d_in, mid, bias=False, dtype=dtype),
                "linear_2": nn.Linear(d_mid, cfg["hidden_size"], bias=False, dtype=dtype)}
        mods["ln_mid" if ln_mid else "ln_post"] = RMSNorm(d_mid if ln_mid else cfg["hidden_size"], 1e-6, dtype)
        self.multi_modal_projector = nn.ModuleDict(mods)
        self.act_name = cfg.get("projector_act", "swiglu")

    def _project(self, x):
        p = self.multi_modal_projector
        T, C = x.shape
        pad = (-T) % self.stack
        x = F.pad(x, (0, 0, 0, pad)).view((T + pad) // self.stack, C * self.stack)
        h = p["linear_1"](p["ln_pre"](x))
        if self.swiglu:
            a, b = h.chunk(2, dim=-1)
            h = a * F.silu(b)
        else:
            h = F.gelu(h)
        if "ln_mid" in p:
            h = p["ln_mid"](h)
        h = p["linear_2"](h)
        return p["ln_post"](h) if "ln_post" in p else h

    @torch.inference_mode()
    def embed_multimodal_item(self, f: dict) -> torch.Tensor:
        mel = torch.as_tensor(f["data"])
        valid = f.get("valid_frames")
        x = self.audio_tower(mel, valid)
        if valid is not None:
            x = x[: (valid - 1) // 2 + 1]
        return self._project(x)[: f["length"]]

    @staticmethod
    def num_tokens(ac: dict, valid_frames: int, stack: int = 8) -> int:
        return max(1, -(-((valid_frames - 1) // 2 + 1) // stack))


class AudioLMProcessor:
    """``multi_modal_data={"audio": clip | [clips]}``: clip = waveform array / (array, sr) / path / bytes, or a precomputed
    log-mel ``[n_mels, frames]`` tensor.  Pads to whole tower windows, records the valid frame count, expands the placeholder."""

    def __init__(self, model_config) -> None:
        from vllm_b200.models.registry import resolve_model_cls
        cfg = model_config.hf_config
        self.ac = cfg["audio_config"]
        self.token = cfg.get("audio_token_index", cfg.get("audio_token_id"))
        self.model_cls = resolve_model_cls(model_config.architecture)
        self.window = 2 * self.ac["max_source_positions"]
        self.multi_window = model_config.architecture.startswith("Voxtral")
        self.stack = int(cfg.get("stack_factor", 8)) if model_config.architecture.startswith("Ultravox") else 0

    def _mel(self, clip: Any) -> tuple[torch.Tensor, int]:
        if isinstance(clip, torch.Tensor) and clip.dim() == 2:
            mel = clip.float()
        else:
            from vllm_b200.multimodal.audio import load_audio, log_mel_spectrogram
            wav = load_audio(clip)
            mel = log_mel_spectrogram(wav, self.ac["num_mel_bins"])
            valid = min(mel.shape[1], len(wav) // 160)
            mel = mel[:, :max(valid, 1)] if not self.multi_window else mel
        valid = mel.shape[1]
        total = self.window * (max(1, -(-valid // self.window)) if self.multi_window else 1)
        if valid > total:
            mel, valid = mel[:, :total], total
        if valid < total:
            mel = F.pad(mel, (0, total - valid), value=float(mel.min()) if mel.numel() else 0.0)
        return mel, valid

    def process(self, ids: list[int], mm_data: dict[str, Any]) -> tuple[list[int], list[dict[str, Any]]]:
        from vllm_b200.multimodal import hash_item
        clips = mm_data.get("audio")
        if clips is None:
            return ids, []
        if not isinstance(clips, (list,)) or (isinstance(clips, tuple) and len(clips) == 2 and isinstance(clips[1], (int, float))):
            clips = [clips]
        slots = [i for i, t in enumerate(ids) if t == self.token]
        if len(slots) != len(clips):
            raise ValueError(f"{len(clips)} audio clips but {len(slots)} audio placeholders in the prompt")
        out, feats, prev = [], [], 0
        for slot, clip in zip(slots, clips):
            mel, valid = self._mel(clip)
            n = int(self.model_cls.num_tokens(self.ac, valid, self.stack) if self.stack else self.model_cls.num_tokens(self.ac, valid))
            out.extend(ids[prev:slot])
            feats.append({"modality": "audio", "hash": hash_item(np.ascontiguousarray(mel.numpy())), "offset": len(out), "length": n,
                          "data": mel, "valid_frames": valid})
            out.extend([self.token] * n)
            prev = slot + 1
        out.extend(ids[prev:])
        return out, feats
