"""Token sampler.

Pipeline parity with the reference's ``Sampler.forward``
(vllm/v1/sample/sampler.py:73: fp32 logits → allowed-ids / bad-words /
min-tokens / logit-bias → penalties → temperature → min-p / top-k / top-p →
Gumbel-max sample → logprobs gather) and MRV2's Gumbel sampler
(worker/gpu/sample/gumbel.py:182).

On CUDA the common cases (greedy, or temperature-only random sampling) run as
one fused sm_100a kernel over the logits (scale + Philox Gumbel noise + argmax,
``ops/cuda.py::fused_sample``); filtered sampling uses a threshold-search
kernel; everything has a PyTorch reference path used on CPU and in tests.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch

from vllm_b200.ops import ref


@dataclass
class SamplingMetadata:
    temperature: torch.Tensor                 # [B] fp32 (0 = greedy)
    all_greedy: bool
    all_random: bool
    top_k: Optional[torch.Tensor] = None      # [B] int32, 0 = off
    top_p: Optional[torch.Tensor] = None      # [B] fp32
    min_p: Optional[torch.Tensor] = None      # [B] fp32
    generators: dict[int, torch.Generator] = field(default_factory=dict)
    max_num_logprobs: Optional[int] = None    # None = no logprobs; 0 = sampled only
    # penalties
    no_penalties: bool = True
    presence: Optional[torch.Tensor] = None
    frequency: Optional[torch.Tensor] = None
    repetition: Optional[torch.Tensor] = None
    prompt_mask: Optional[torch.Tensor] = None    # [B,V] bool
    output_counts: Optional[torch.Tensor] = None  # [B,V] int32
    # logits processors
    logit_bias: list[Optional[dict[int, float]]] = field(default_factory=list)
    allowed_token_ids_mask: Optional[torch.Tensor] = None  # [B,V] bool, True = disallowed
    bad_words_token_ids: dict[int, list[list[int]]] = field(default_factory=dict)
    min_tokens: dict[int, tuple[int, set[int]]] = field(default_factory=dict)  # row → (min, stop ids)
    output_token_ids: list[list[int]] = field(default_factory=list)
    extra_args: list[Optional[dict]] = field(default_factory=list)   # per row ``SamplingParams.extra_args`` (custom processors)
    prompt_token_ids: list[list[int]] = field(default_factory=list)
    seeds: Optional[torch.Tensor] = None      # [B] int64 per-request noise seeds (CUDA path)
    steps: Optional[torch.Tensor] = None      # [B] int32 per-request noise counter = #output tokens so far


@dataclass
class SamplerOutput:
    sampled_token_ids: torch.Tensor                      # [B] int64 (or [B, k] for spec decode)
    logprobs: Optional[tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None  # ids, vals, ranks


class LogitsProcessor:
    """Plug-in interface (``--logits-processors pkg.module:Class``; reference: v1/sample/logits_processor/interface.py).
    One instance per engine; ``apply`` sees the whole batch and edits rows whose request opted in through
    ``SamplingParams.extra_args`` (requests with extra_args take the per-request sampling path)."""

    def __init__(self, vllm_config, device) -> None:
        self.vllm_config, self.device = vllm_config, device

    def is_argmax_invariant(self) -> bool:
        """True if the processor can never change which token has the highest logit (then greedy rows skip it)."""
        return False

    def apply(self, logits: torch.Tensor, md: "SamplingMetadata") -> torch.Tensor:
        raise NotImplementedError


def load_logits_processors(specs, vllm_config, device) -> list[LogitsProcessor]:
    from vllm_b200.utils import resolve_obj_by_qualname
    out = []
    for spec in specs or ():
        cls = resolve_obj_by_qualname(spec.replace(":", ".")) if isinstance(spec, str) else spec
        out.append(cls(vllm_config, device))
    return out


class Sampler:
    def __init__(self, logits_processors: Optional[list] = None) -> None:
        self.custom = list(logits_processors or [])

    def __call__(self, logits: torch.Tensor, md: SamplingMetadata,
                 grammar_bitmask: Optional[torch.Tensor] = None) -> SamplerOutput:
        logits = logits.float()
        raw_for_logprobs = logits if md.max_num_logprobs is not None else None
        if grammar_bitmask is not None:
            logits = apply_bitmask(logits, grammar_bitmask)
        logits = self.apply_logits_processors(logits, md)
        for proc in self.custom:
            if md.extra_args and any(md.extra_args) and not (md.all_greedy and proc.is_argmax_invariant()):
                logits = proc.apply(logits, md)
        if not md.no_penalties:
            logits = ref.apply_penalties(logits, md.prompt_mask, md.output_counts, md.presence, md.frequency,
                                         md.repetition)
        sampled = self.sample(logits, md)
        lp = None
        if md.max_num_logprobs is not None:
            lp = ref.compute_logprobs(raw_for_logprobs, sampled, md.max_num_logprobs)
        return SamplerOutput(sampled, lp)

    @staticmethod
    def apply_logits_processors(logits: torch.Tensor, md: SamplingMetadata) -> torch.Tensor:
        if md.allowed_token_ids_mask is not None:
            logits = logits.masked_fill(md.allowed_token_ids_mask, float("-inf"))
        for row, (mn, stop_ids) in md.min_tokens.items():
            if len(md.output_token_ids[row]) < mn and stop_ids:
                logits[row, list(stop_ids)] = float("-inf")
        for row, bias in enumerate(md.logit_bias):
            if bias:
                idx = torch.tensor(list(bias.keys()), device=logits.device, dtype=torch.long)
                val = torch.tensor(list(bias.values()), device=logits.device, dtype=logits.dtype)
                logits[row, idx] += val
        for row, seqs in md.bad_words_token_ids.items():
            past = md.output_token_ids[row]
            for seq in seqs:
                if len(seq) == 1 or (len(seq) - 1 <= len(past) and past[len(past) - len(seq) + 1:] == seq[:-1]):
                    logits[row, seq[-1]] = float("-inf")
        return logits

    def sample(self, logits: torch.Tensor, md: SamplingMetadata) -> torch.Tensor:
        if md.all_greedy:
            return greedy(logits)
        if logits.is_cuda and md.top_k is None and md.top_p is None and md.min_p is None \
                and md.seeds is not None:
            from vllm_b200.ops import cuda_ops
            return cuda_ops().fused_sample(logits, md.temperature, md.seeds, md.steps)
        filtered = logits
        if md.top_k is not None or md.top_p is not None or md.min_p is not None:
            t = torch.where(md.temperature < 1e-5, torch.ones_like(md.temperature), md.temperature)
            if logits.is_cuda and md.seeds is not None and logits.dtype == torch.float32:
                # device path: exact top-k / nucleus / min-p masking by integer-key bisection (no sort), then the
                # fused Gumbel-max sampler (temperature already applied → 1, greedy rows stay greedy with 0)
                from vllm_b200.ops import cuda_ops
                cu = cuda_ops()
                scaled = (logits / t[:, None]).contiguous()
                cu.topk_topp_mask_(scaled, md.top_k, md.top_p, md.min_p)
                t_eff = torch.where(md.temperature < 1e-5, md.temperature, torch.ones_like(md.temperature))
                return cu.fused_sample(scaled, t_eff, md.seeds, md.steps)
            filtered = ref.apply_top_k_top_p_min_p(logits / t[:, None], md.top_k, md.top_p, md.min_p)
            return ref.gumbel_sample(filtered, torch.where(md.temperature < 1e-5, md.temperature,
                                                           torch.ones_like(md.temperature)), md.generators)
        return ref.gumbel_sample(filtered, md.temperature, md.generators)


def greedy(logits: torch.Tensor) -> torch.Tensor:
    if logits.is_cuda:
        from vllm_b200.ops import cuda_ops
        return cuda_ops().argmax(logits)
    return logits.argmax(dim=-1)


def apply_bitmask(logits: torch.Tensor, bitmask: torch.Tensor) -> torch.Tensor:
    """bitmask: [B, ceil(V/32)] int32, bit set = token allowed (xgrammar convention)."""
    V = logits.shape[-1]
    bits = torch.arange(32, device=bitmask.device, dtype=torch.int32)
    allowed = ((bitmask[:, :, None] >> bits) & 1).bool().reshape(bitmask.shape[0], -1)[:, :V]
    return logits.masked_fill(~allowed, float("-inf"))


class RejectionSampler:
    """Speculative-decoding verification (reference: vllm/v1/sample/rejection_sampler.py:38).

    Greedy: accept draft token i while it equals the target argmax; emit the
    first mismatch's target token (or the bonus token if all accepted).
    Random: standard accept with prob min(1, p/q) for draft-model proposals;
    n-gram drafts have q = 1 (deterministic proposal).

    ``md`` is EXPANDED: one metadata row per target-logit row (request i, draft position j → row cu[i] + j, whose
    ``output_token_ids`` already contain drafts[:j]), so that the same logits processors, grammar bitmask, penalties
    and top-k / top-p / min-p filters as the non-speculative ``Sampler`` shape every verified position — a speculated
    run then samples from exactly the distribution the plain run would (reference: rejection_sampler.py
    ``apply_logits_processors`` / ``apply_sampling_constraints``)."""

    def __init__(self, sampler: Optional["Sampler"] = None) -> None:
        self.sampler = sampler or Sampler()

    def process(self, logits: torch.Tensor, md: SamplingMetadata, bitmask: Optional[torch.Tensor]) -> torch.Tensor:
        """Raw target logits → the logits the plain sampler would draw from (temperature applied on random rows)."""
        logits = logits.float()
        if bitmask is not None:
            logits = apply_bitmask(logits, bitmask)
        logits = Sampler.apply_logits_processors(logits, md)
        for proc in self.sampler.custom:
            if md.extra_args and any(md.extra_args) and not (md.all_greedy and proc.is_argmax_invariant()):
                logits = proc.apply(logits, md)
        if not md.no_penalties:
            logits = ref.apply_penalties(logits, md.prompt_mask, md.output_counts, md.presence, md.frequency,
                                         md.repetition)
        if not md.all_greedy:
            t = torch.where(md.temperature < 1e-5, torch.ones_like(md.temperature), md.temperature)
            logits = logits / t[:, None]
            if md.top_k is not None or md.top_p is not None or md.min_p is not None:
                logits = ref.apply_top_k_top_p_min_p(logits, md.top_k, md.top_p, md.min_p)
        return logits

    def __call__(self, target_logits: torch.Tensor, draft_ids: list[list[int]], md: SamplingMetadata,
                 cu: list[int], bitmask: Optional[torch.Tensor] = None) -> list[list[int]]:
        """target_logits: [sum(k_i + 1), V] — for request i rows cu[i]:cu[i+1] are the logits
        after each of its k_i drafts' prefixes (+1 bonus)."""
        out: list[list[int]] = []
        logits = self.process(target_logits, md, bitmask)
        tgt = logits.argmax(dim=-1).tolist() if md.all_greedy else None
        temps = md.temperature.tolist() if not md.all_greedy else None
        for i, drafts in enumerate(draft_ids):
            s = cu[i]
            acc: list[int] = []
            if md.all_greedy or temps[s] < 1e-5:
                row_t = tgt[s:s + len(drafts) + 1] if tgt is not None else \
                    logits[s:s + len(drafts) + 1].argmax(-1).tolist()
                for j, d in enumerate(drafts):
                    if row_t[j] == d:
                        acc.append(d)
                    else:
                        break
                acc.append(row_t[len(acc)])
            else:
                g = md.generators.get(s)
                probs = logits[s:s + len(drafts) + 1].softmax(-1)
                ok = True
                for j, d in enumerate(drafts):
                    u = torch.rand((), generator=g, device=probs.device if g is None else g.device).item()
                    if u < probs[j, d].item():  # q(d) = 1 for deterministic drafts
                        acc.append(d)
                    else:
                        p = probs[j].clone()
                        p[d] = 0
                        tot = p.sum()
                        if float(tot) <= 0:      # the draft was the only admissible token and lost the coin flip
                            acc.append(d)
                            continue
                        acc.append(int(torch.multinomial(p / tot, 1, generator=g).item()))
                        ok = False
                        break
                if ok:
                    acc.append(int(torch.multinomial(probs[len(drafts)], 1, generator=g).item()))
            out.append(acc)
        return out
