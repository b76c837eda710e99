"""EAGLE / MTP style proposers: a one-layer draft head that reads the TARGET model's hidden states.

Behavioural reference: vllm/v1/spec_decode/eagle.py (``EagleProposer``: first pass over the newly verified
positions with (next-token embedding, target hidden) pairs, then k-1 autoregressive passes on its own hidden
output), model_executor/models/llama_eagle.py, deepseek_mtp.py.

Slot t of the head consumes ``fc([embed(x_{t+1}); h_t])`` (``h_t`` = target hidden state at position t) and predicts
``x_{t+2}``.  After every target step the head first catches up over the positions the target just produced
(writing its own KV), which yields the first draft; further drafts feed its own output hidden state back in.
The head owns a private position-indexed KV arena (one layer), so the engine's block manager is not involved and
rejected drafts are simply overwritten.  ``method="mtp"`` adds the DeepSeek multi-token-prediction input norms.
"""
from __future__ import annotations

import os
from typing import Iterable, Optional

import numpy as np
import torch
import torch.nn as nn

from vllm_b200.forward_context import AttentionMetadata, ForwardContext, set_forward_context
from vllm_b200.layers.common import RMSNorm
from vllm_b200.logger import init_logger
from vllm_b200.models.llama import LlamaDecoderLayer
from vllm_b200.utils import cdiv

logger = init_logger(__name__)


class EagleHead(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, mtp: bool, prefix: str = "eagle.", layer_factory=None) -> None:
        super().__init__()
        H = cfg["hidden_size"]
        eps = cfg.get("rms_norm_eps", 1e-6)
        self.mtp = mtp
        self.enorm = RMSNorm(H, eps, dtype) if mtp else None
        self.hnorm = RMSNorm(H, eps, dtype) if mtp else None
        self.fc = nn.Linear(2 * H, H, bias=False, dtype=dtype)
        # MTP heads are one more decoder layer OF THE TARGET'S OWN KIND (MLA + experts for DeepSeek-V3, q/k-normed GQA +
        # experts for GLM-4.5, …): the target model's layer factory builds it; EAGLE heads are a Llama layer
        layer = layer_factory(f"{prefix}layers.0") if layer_factory is not None else \
            LlamaDecoderLayer(cfg, mc, dtype, None, f"{prefix}layers.0", 0)
        self.layers = nn.ModuleList([layer])
        self.native_layer = layer_factory is not None
        self.norm = RMSNorm(H, eps, dtype)

    def forward(self, emb: torch.Tensor, hidden: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        if self.mtp:
            emb, hidden = self.enorm(emb), self.hnorm(hidden)
        x = self.fc(torch.cat((emb, hidden), dim=-1))
        layer = self.layers[0]
        out, _ = layer(positions, layer.input_layernorm(x), x, self.norm)
        return out

    stacked = [("qkv_proj", "q_proj", "q"), ("qkv_proj", "k_proj", "k"), ("qkv_proj", "v_proj", "v"),
               ("gate_up_proj", "gate_proj", 0), ("gate_up_proj", "up_proj", 1)]

    # checkpoint names of the MTP block inside a TARGET checkpoint → ours (DeepSeek-V3 / GLM-4.5 / ERNIE: extra decoder
    # layer index L; Qwen3-Next: ``mtp.*``; MiMo: ``model.mtp_layers.0.*``)
    _MTP_RENAMES = [("eh_proj.", "fc."), ("shared_head.norm.", "norm."), ("pre_fc_norm_embedding.", "enorm."),
                    ("pre_fc_norm_hidden.", "hnorm."), ("token_layernorm.", "enorm."), ("hidden_layernorm.", "hnorm."),
                    ("input_proj.", "fc."), ("final_layernorm.", "norm."), ("mtp_block.", "layers.0.")]
    _EXPERT = {"gate_proj": ("w13_weight", "w1"), "up_proj": ("w13_weight", "w3"), "down_proj": ("w2_weight", "w2"),
               "w1": ("w13_weight", "w1"), "w3": ("w13_weight", "w3"), "w2": ("w2_weight", "w2")}

    def load_mtp_weights(self, weights: Iterable[tuple[str, torch.Tensor]], num_layers: int) -> int:
        """Pick the multi-token-prediction tensors out of the target's checkpoint stream and load them."""
        pres = (f"model.layers.{num_layers}.", "mtp.layers.0.", "mtp.", "model.mtp_layers.0.", "model.mtp.layers.0.", "model.mtp.")
        top = ("enorm.", "hnorm.", "fc.", "norm.")

        def picked():
            for name, w in weights:
                pre = next((p for p in pres if name.startswith(p)), None)
                if pre is None:
                    continue
                rest = name[len(pre):]
                if rest.startswith(("embed_tokens.", "shared_head.head.")):        # shared with the target
                    continue
                for a, b in self._MTP_RENAMES:
                    if rest.startswith(a):
                        rest = b + rest[len(a):]
                yield (rest if rest.startswith(top) or rest.startswith("layers.0.") else "layers.0." + rest), w
        return self.load_weights(picked())

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> int:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        n = 0
        for name, w in weights:
            name = name.removeprefix("model.")
            if ".mlp.experts." in name and name.split(".mlp.experts.")[1].split(".")[0].isdigit():
                head, tail = name.split(".mlp.experts.", 1)
                e, which = tail.split(".")[0], tail.split(".")[1]
                suffix, shard = self._EXPERT[which]
                pn = f"{head}.mlp.experts.{suffix}"
                if pn in params:
                    modules[f"{head}.mlp.experts"].weight_loader(params[pn], w, shard, int(e))
                    n += 1
                continue
            name = name.replace(".mlp.gate.e_score_correction_bias", ".mlp.experts.e_score_correction_bias")
            target, shard = name, None
            for fused, part, sid in self.stacked:
                if f".{part}." in name:
                    target, shard = name.replace(part, fused), sid
                    break
            if target not in params:
                continue
            mod = modules[target.rsplit(".", 1)[0]]
            loader = getattr(mod, "weight_loader", None)
            loader(params[target], w, shard) if loader is not None else params[target].data.copy_(w)
            n += 1
        return n


class Eagle3Layer(nn.Module):
    """EAGLE-3 draft layer: attention reads [norm(token embedding) ; norm(hidden)] (2H wide), the residual stream
    carries the hidden half only (reference: model_executor/models/llama_eagle3.py)."""

    def __init__(self, cfg: dict, mc, dtype, prefix: str) -> None:
        super().__init__()
        from vllm_b200.layers.attention import Attention
        from vllm_b200.layers.linear import QKVParallelLinear, RowParallelLinear
        from vllm_b200.layers.rotary import get_rope
        from vllm_b200.models.llama import LlamaMLP
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        nkv = cfg.get("num_key_value_heads") or nh
        hd = cfg.get("head_dim") or H // nh
        eps = cfg.get("rms_norm_eps", 1e-6)
        self.qkv_proj = QKVParallelLinear(2 * H, hd, nh, nkv, False, dtype, None, f"{prefix}.self_attn.qkv_proj")
        self.o_proj = RowParallelLinear(nh * hd, H, False, dtype=dtype, prefix=f"{prefix}.self_attn.o_proj")
        rope = get_rope(hd, hd, mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 8192),
                        cfg.get("rope_theta", 10000.0), True, cfg.get("rope_scaling"))
        self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads,
                              prefix=f"{prefix}.self_attn.attn", rope=rope)
        self.mlp = LlamaMLP(H, cfg["intermediate_size"], cfg.get("hidden_act", "silu"), False, dtype, None, f"{prefix}.mlp")
        self.input_layernorm, self.hidden_norm = RMSNorm(H, eps, dtype), RMSNorm(H, eps, dtype)
        self.post_attention_layernorm = RMSNorm(H, eps, dtype)
        self.sizes = [self.qkv_proj.q_size, self.qkv_proj.kv_size, self.qkv_proj.kv_size]

    def forward(self, positions, emb, hidden):
        x = torch.cat((self.input_layernorm(emb), self.hidden_norm(hidden)), dim=-1)
        q, k, v = self.qkv_proj(x).split(self.sizes, dim=-1)
        h = hidden + self.o_proj(self.attn(q, k, v, positions))
        return h + self.mlp(self.post_attention_layernorm(h))


class Eagle3Head(nn.Module):
    """fc(3H → H) over the tapped target layers, one Eagle3Layer, its own norm; drafts are read through the target's
    LM head (a reduced draft vocabulary + ``d2t`` table is applied when the checkpoint ships one)."""

    def __init__(self, cfg: dict, mc, dtype, n_taps: int, prefix: str = "eagle3.") -> None:
        super().__init__()
        H = cfg["hidden_size"]
        self.H = H
        self.fc = nn.Linear(n_taps * H, H, bias=False, dtype=dtype)
        self.layers = nn.ModuleList([Eagle3Layer(cfg, mc, dtype, f"{prefix}layers.0")])
        self.norm = RMSNorm(H, cfg.get("rms_norm_eps", 1e-6), dtype)
        self.d2t: Optional[torch.Tensor] = None

    def forward(self, emb: torch.Tensor, hidden: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        if hidden.shape[-1] != self.H:               # first pass: concatenated taps; later passes feed our own state back
            hidden = self.fc(hidden)
        return self.layers[0](positions, emb, hidden)

    def readout(self, h: torch.Tensor) -> torch.Tensor:
        return self.norm(h)

    load_weights = EagleHead.load_weights
    stacked = EagleHead.stacked


class EagleDrafter:
    needs_step_hidden = True
    aux_layers: Optional[tuple] = None

    def __init__(self, vllm_config, runner) -> None:
        from vllm_b200.layers import attention as attn_mod
        sc = vllm_config.speculative_config
        mc = vllm_config.model_config
        self.k = sc.num_speculative_tokens
        self.device = runner.device
        before = set(attn_mod.attention_layers())
        old = torch.get_default_dtype()
        torch.set_default_dtype(mc.torch_dtype)
        try:
            with torch.device(self.device):
                if sc.method == "eagle3":
                    L = mc.hf_config["num_hidden_layers"]
                    self.aux_layers = tuple(sorted({min(2, L - 1), L // 2, max(L - 3, 0)}))     # low / mid / high taps
                    self.head = Eagle3Head(mc.hf_config, mc, mc.torch_dtype, len(self.aux_layers))
                else:
                    self.head = EagleHead(mc.hf_config, mc, mc.torch_dtype, mtp=sc.method == "mtp",
                                          layer_factory=self._native_layer_factory(runner, mc) if sc.method == "mtp" else None)
        finally:
            torch.set_default_dtype(old)
        reg = attn_mod.attention_layers()
        self.layers = {n: reg.pop(n) for n in list(reg) if n not in before}
        with torch.no_grad():
            for p in self.head.parameters():
                if p.dim() > 1:
                    p.uniform_(-0.05, 0.05)
        if sc.model and os.path.isdir(sc.model):
            from vllm_b200.model_loader import iter_weights
            logger.info("EAGLE head: loaded %d tensors from %s", self.head.load_weights(iter_weights(sc.model, "auto")),
                        sc.model)
        elif sc.method == "mtp" and os.path.isdir(mc.model):
            # the prediction block ships inside the target checkpoint
            from vllm_b200.model_loader import iter_weights
            got = self.head.load_mtp_weights(iter_weights(mc.model, "auto"), mc.hf_config["num_hidden_layers"])
            logger.info("MTP head (%s layer): loaded %d tensors from the target checkpoint",
                        "native" if self.head.native_layer else "Llama", got)
        self.head.eval()
        self.bs = vllm_config.cache_config.block_size
        R = runner.max_num_reqs
        self.blocks_per_row = cdiv(runner.max_model_len + self.k + 1, self.bs)
        self.kv = {}
        from vllm_b200.core.kv_cache_manager import MambaSpec, MLAAttentionSpec
        for name, layer in self.layers.items():
            spec = layer.get_kv_cache_spec(self.bs, mc.torch_dtype.itemsize) if hasattr(layer, "get_kv_cache_spec") else None
            if isinstance(spec, MambaSpec):
                raise NotImplementedError("multi-token-prediction blocks with recurrent (state-space / delta-net) mixers")
            shape = ((R * self.blocks_per_row, self.bs, layer.head_size) if isinstance(spec, MLAAttentionSpec) else
                     (R * self.blocks_per_row, 2, layer.num_kv_heads, self.bs, layer.head_size))
            self.kv[name] = torch.zeros(shape, dtype=mc.torch_dtype, device=self.device)
        self.block_table = torch.arange(R * self.blocks_per_row, dtype=torch.int32, device=self.device).view(R, -1)
        self.layer_to_group = {n: 0 for n in self.layers}
        self.row_req: dict[int, str] = {}          # arena row → request id it currently serves
        self.base = np.zeros(R, dtype=np.int64)    # first position the head saw for that request (prefix-cache hits)
        self.filled = np.zeros(R, dtype=np.int64)  # positions < filled have valid head KV

    @staticmethod
    def _native_layer_factory(runner, mc):
        """A builder for one more decoder layer of the target's own kind (index = num_hidden_layers), or None when the
        target model does not expose its layer factory (then the head falls back to a Llama layer)."""
        inner = getattr(runner.model, "model", None)
        if inner is None and hasattr(runner.model, "language_model"):
            inner = getattr(runner.model.language_model, "model", None)
        make = getattr(inner, "_make_layer", None)
        if make is None:
            return None
        cfg = mc.hf_config
        L = cfg["num_hidden_layers"]
        # per-layer lists are indexed by the layer id: give index L the kind the checkpoints use for the prediction block
        for key, fill in (("layer_types", "full_attention"), ("mlp_layer_types", "sparse"), ("no_rope_layers", 1)):
            v = cfg.get(key)
            if isinstance(v, list) and len(v) == L:
                v.append(fill)
        return lambda prefix: make(cfg, mc, mc.torch_dtype, None, prefix, L)

    def _run(self, emb, hidden, pos, md, last_idx):
        ctx = ForwardContext(attn_metadata={0: md}, kv_caches=self.kv, layer_to_group=self.layer_to_group,
                             num_tokens=emb.shape[0])
        with set_forward_context(ctx):
            out = self.head(emb, hidden, pos)
        return out[last_idx]

    @torch.inference_mode()
    def propose(self, runner, rows, sampled: list[list[int]]) -> list[list[int]]:
        out: list[list[int]] = [[] for _ in rows]
        ctx = getattr(runner, "step_hidden", None)
        if ctx is None:
            return out
        hidden, h_rows, cu, c0s = ctx
        runner.step_hidden = None
        pos_of = {int(r): j for j, r in enumerate(h_rows)}
        dev, bs = self.device, self.bs
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)  # noqa: E731
        seg_tok, seg_pos, seg_hid, seg_slots, lens, arows, want = [], [], [], [], [], [], []
        for j, r in enumerate(rows):
            r = int(r)
            hj = pos_of.get(r)
            if hj is None:
                continue
            rid = runner.req_ids[r]
            c0 = int(c0s[hj])
            c1 = int(runner.num_computed_cpu[r])              # target KV (and hidden) valid for [c0, c1)
            ntok = int(runner.num_tokens_cpu[r])
            c1 = min(c1, ntok - 1)                            # slot t needs token t+1
            if self.row_req.get(r) != rid or c0 < self.filled[r] and c0 < self.base[r]:
                self.row_req[r], self.base[r], self.filled[r] = rid, c0, c0
            c0 = max(c0, int(self.base[r]))
            if c1 <= c0:
                continue
            idx = np.arange(c0, c1)
            seg_tok.append(runner.token_ids_cpu[r, idx + 1].astype(np.int64))
            seg_pos.append(idx.astype(np.int64))
            seg_hid.append(hidden[int(cu[hj]) + (c0 - int(c0s[hj])): int(cu[hj]) + (c1 - int(c0s[hj]))])
            seg_slots.append(r * self.blocks_per_row * bs + (idx - int(self.base[r])))
            lens.append(c1 - c0)
            arows.append(r)
            self.filled[r] = c1
            room = runner.max_model_len - ntok - 1
            want.append((j, min(self.k, max(room, 0))) if sampled[j] else (j, 0))
        if not arows:
            return out
        B = len(arows)
        lens_a = np.array(lens, dtype=np.int64)
        qsl = np.zeros(B + 1, dtype=np.int32)
        qsl[1:] = np.cumsum(lens_a)
        seqs = np.array([self.filled[r] - self.base[r] for r in arows], dtype=np.int32)
        bt = self.block_table[t(np.array(arows), torch.int64)]
        md = AttentionMetadata(num_reqs=B, num_actual_tokens=int(qsl[-1]), max_query_len=int(lens_a.max()),
                               max_seq_len=int(seqs.max()), query_start_loc=t(qsl, torch.int32),
                               seq_lens=t(seqs, torch.int32), block_table=bt,
                               slot_mapping=t(np.concatenate(seg_slots), torch.int64),
                               query_start_loc_cpu=torch.from_numpy(qsl), seq_lens_cpu=torch.from_numpy(seqs))
        emb = runner.model.embed_input_ids(t(np.concatenate(seg_tok), torch.int64))
        last = t(qsl[1:] - 1, torch.int64)
        h = self._run(emb, torch.cat(seg_hid, dim=0), t(np.concatenate(seg_pos), torch.int64), md, last)
        if not any(k for _, k in want):
            return out
        drafts = []
        one = torch.arange(B + 1, dtype=torch.int32, device=dev)
        for step in range(self.k):
            hn = self.head.readout(h) if hasattr(self.head, "readout") else h
            tok = runner.model.compute_logits(hn).float().argmax(-1)
            drafts.append(tok)
            if step == self.k - 1:
                break
            p = np.array([self.filled[r] + step for r in arows], dtype=np.int64)      # slot being written
            sl = np.array([r * self.blocks_per_row * bs + (self.filled[r] + step - self.base[r]) for r in arows])
            cur = (p - np.array([self.base[r] for r in arows]) + 1).astype(np.int32)
            md = AttentionMetadata(num_reqs=B, num_actual_tokens=B, max_query_len=1, max_seq_len=int(cur.max()),
                                   query_start_loc=one, seq_lens=t(cur, torch.int32), block_table=bt,
                                   slot_mapping=t(sl, torch.int64),
                                   query_start_loc_cpu=torch.arange(B + 1, dtype=torch.int32),
                                   seq_lens_cpu=torch.from_numpy(cur), decode_only=True)
            h = self._run(runner.model.embed_input_ids(tok), h, t(p, torch.int64), md, torch.arange(B, device=dev))
        mat = torch.stack(drafts, dim=1).cpu().numpy()
        for b, (j, k) in enumerate(want):
            out[j] = mat[b, :k].tolist()
        return out
