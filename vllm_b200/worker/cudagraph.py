"""Full-step CUDA graphs for decode batches.

Reference: vllm/v1/worker/gpu/cudagraph_utils.py:102 (``CudaGraphManager``), config/compilation.py:53-64
(``CUDAGraphMode``).  One graph per padded batch bucket captures forward + LM head + fused sampling over
the runner's static buffers; everything variable (sequence lengths, block tables, slot mapping,
temperatures, seeds, step counter) lives in device buffers that are overwritten before replay, and the
attention launch geometry depends only on the bucket, never on sequence lengths.
"""
from __future__ import annotations

import torch

from vllm_b200.forward_context import ForwardContext, set_forward_context
from vllm_b200.layers.attention import attention_layers
from vllm_b200.ops import cuda_ops


class DecodeGraph:
    def __init__(self, runner, b: int) -> None:
        self.runner = runner
        self.b = b
        dev = runner.device
        self.temperature = torch.zeros(b, dtype=torch.float32, device=dev)
        self.seeds = torch.zeros(b, dtype=torch.int64, device=dev)
        self.steps = torch.zeros(b, dtype=torch.int32, device=dev)
        self.out = torch.zeros(b, dtype=torch.int64, device=dev)
        # safe placeholder inputs for warm-up / capture: one token at position 0 of the null block
        runner.input_ids[:b].zero_()
        runner.positions[:b].zero_()
        runner.seq_lens[:b].fill_(1)
        runner.qsl[:b + 1].copy_(torch.arange(b + 1, dtype=torch.int32, device=dev))
        for g in range(runner.num_kv_groups):
            runner.slot_mappings[g][:b].fill_(-1)
            runner.block_tables[g][:b].zero_()
        self.md = runner._make_metadata(b, b, 1, runner.max_model_len, True)
        layers = list(attention_layers().values())
        hkv = max((getattr(l, "num_kv_heads", 1) for l in layers), default=1)
        for m in self.md.values():
            # typical context assumed for the split heuristic; the kernel itself adapts to real lengths
            m.num_splits = cuda_ops().decode_splits(b, hkv, min(runner.max_model_len, 4096),
                                                    128 if m.kv_vt else runner.block_size)
        self.graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if runner.graph_pool is None:
            runner.graph_pool = torch.cuda.graph_pool_handle()
        n0 = cuda_ops().LAUNCHES
        with torch.cuda.graph(self.graph, pool=runner.graph_pool):
            self._body()
        self.num_launches = cuda_ops().LAUNCHES - n0

    def _body(self) -> None:
        r, b = self.runner, self.b
        ctx = ForwardContext(attn_metadata=self.md, kv_caches=r.kv_caches, layer_to_group=r.layer_to_group,
                             num_tokens=b)
        with set_forward_context(ctx):
            hidden = r.model(r.input_ids[:b], r.positions[:b], None)
        fast = getattr(r.model, "sample_candidates", None)
        if fast is not None and fast(hidden, self.temperature, self.seeds, self.steps, self.out):
            return                      # TP: per-shard sampling + 16-byte candidate gather instead of the logits all-gather
        logits = r.model.compute_logits(hidden)
        cuda_ops().fused_sample(logits, self.temperature, self.seeds, self.steps, self.out)

    def replay_from_staging(self, stg, b: int) -> torch.Tensor:
        """Sampling vectors come from the pinned staging set (filled from the runner's per-row arrays)."""
        self.temperature.copy_(stg.temperature[:b], non_blocking=True)
        self.seeds.copy_(stg.seeds[:b], non_blocking=True)
        self.steps.copy_(stg.steps[:b], non_blocking=True)
        self.graph.replay()
        self.replays = getattr(self, "replays", 0) + 1
        return self.out


class MixedGraph(DecodeGraph):
    """Full-step graph for a MIXED batch (prefill chunks + decodes) of at most ``t`` tokens in at most ``b`` requests.

    Reference: vllm/v1/cudagraph_dispatcher.py (padded buckets keyed by token count), config/compilation.py:53-64
    (FULL / FULL_AND_PIECEWISE).  The reference captures the token-wise pieces and leaves attention eager; here the
    whole step is one graph because nothing in it depends on the batch composition on the host side: the input expansion
    runs on the device (csrc/input_prep.cu: empty padding requests, inert padding tokens), the varlen attention kernel
    derives its (request, query tile) schedule from ``query_start_loc`` on the device (``AttentionMetadata.flat``), and
    the rows that feed the LM head are gathered with device-side indices."""

    def __init__(self, runner, t: int, b: int) -> None:
        self.t = t
        self.runner = runner
        self.b = b
        dev = runner.device
        self.temperature = torch.zeros(b, dtype=torch.float32, device=dev)
        self.seeds = torch.zeros(b, dtype=torch.int64, device=dev)
        self.steps = torch.zeros(b, dtype=torch.int32, device=dev)
        self.out = torch.zeros(b, dtype=torch.int64, device=dev)
        # placeholder inputs for warm-up / capture: one request of t tokens at positions 0..t-1 writing nowhere
        runner.input_ids[:t].zero_()
        runner.positions[:t].copy_(torch.arange(t, dtype=torch.int64, device=dev))
        runner.seq_lens[:b].fill_(1)
        runner.seq_lens[:1].fill_(t)
        runner.qsl[:b + 1].fill_(t)
        runner.qsl[:1].zero_()
        for g in range(runner.num_kv_groups):
            runner.slot_mappings[g][:t].fill_(-1)
            runner.block_tables[g][:b].zero_()
        self.md = runner._make_metadata(b, t, t, runner.max_model_len, False)
        for m in self.md.values():
            m.flat = True
        self.graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if runner.graph_pool is None:
            runner.graph_pool = torch.cuda.graph_pool_handle()
        n0 = cuda_ops().LAUNCHES
        with torch.cuda.graph(self.graph, pool=runner.graph_pool):
            self._body()
        self.num_launches = cuda_ops().LAUNCHES - n0

    def _body(self) -> None:
        r, b, t = self.runner, self.b, self.t
        ctx = ForwardContext(attn_metadata=self.md, kv_caches=r.kv_caches, layer_to_group=r.layer_to_group, num_tokens=t)
        with set_forward_context(ctx):
            hidden = r.model(r.input_ids[:t], r.positions[:t], None)
        # last scheduled token of every request (empty padding requests resolve to the batch's last real token)
        last = hidden.index_select(0, (r.qsl[1:b + 1] - 1).clamp_min(0).to(torch.int64))
        fast = getattr(r.model, "sample_candidates", None)
        if fast is not None and fast(last, self.temperature, self.seeds, self.steps, self.out):
            return
        logits = r.model.compute_logits(last)
        cuda_ops().fused_sample(logits, self.temperature, self.seeds, self.steps, self.out)
