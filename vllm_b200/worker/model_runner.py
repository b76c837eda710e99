"""Model runner: persistent batch, input preparation, CUDA-graph decode, sampling.

Behavioural reference: the reference's MRV2 runner
(vllm/v1/worker/gpu/model_runner.py:1397 ``execute_model``, ``sample_tokens``
:1682, ``capture_model`` :840, ``AsyncOutput`` async_utils.py:115) and V1
``gpu_input_batch.py`` persistent batch.

Design (ours):
* request state lives in NumPy arrays indexed by a dense row; per-step inputs
  (ids, positions, slot mapping, seq lens, block table) are produced with
  vectorised NumPy into double-buffered *pinned* staging buffers and copied with
  async H2D copies into *static* device buffers — the buffers CUDA graphs replay on;
* decode-only steps replay a full-step CUDA graph per padded batch bucket
  (forward + logits + fused sampling), everything else runs eagerly;
* sampled ids stay on the device in ``last_sampled`` (indexed by row); with async
  scheduling the next step's input ids are patched GPU-side from it, so the host
  never waits for step N before launching step N+1 — the D2H copy of the ids goes
  through a pinned buffer + event and is only awaited when the engine consumes
  the step's output (``AsyncStepOutput.get``).
"""
from __future__ import annotations

import os

import time
import zlib
from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np
import torch

from vllm_b200.config import VllmConfig
from vllm_b200.core.kv_cache_manager import KVCacheConfig, KVCacheSpec
from vllm_b200.core.sched_output import LogprobsTensors, ModelRunnerOutput, SchedulerOutput
from vllm_b200.forward_context import AttentionMetadata, ForwardContext, set_forward_context
from vllm_b200.layers.attention import attention_layers, clear_attention_layers
from vllm_b200.logger import init_logger
from vllm_b200.sample.sampler import RejectionSampler, Sampler, SamplingMetadata
from vllm_b200.utils import cdiv

logger = init_logger(__name__)


@dataclass
class ReqState:
    req_id: str
    row: int
    sampling_params: Any
    pooling_params: Any
    generator: Optional[torch.Generator]
    lora_request: Any = None
    num_prompt_tokens: int = 0
    output_token_ids: list[int] = field(default_factory=list)
    mm_features: Optional[list] = None
    complex_sampling: bool = False  # needs the per-request slow path (penalties, logprobs, processors, ...)


class _Staging:
    """One set of pinned host buffers (inputs H2D, sampled ids D2H) + the event that guards their reuse."""

    def __init__(self, T: int, R: int, MB: int, G: int, pin: bool) -> None:
        def mk(*shape, dtype):
            t = torch.zeros(*shape, dtype=dtype)
            return t.pin_memory() if pin else t
        self.input_ids = mk(T, dtype=torch.int32)
        self.positions = mk(T, dtype=torch.int64)
        self.qsl = mk(R + 1, dtype=torch.int32)
        self.seq_lens = mk(R, dtype=torch.int32)
        self.slots = [mk(T, dtype=torch.int64) for _ in range(G)]
        self.block_tables = [mk(R, MB, dtype=torch.int32) for _ in range(G)]
        self.sampled = mk(R, dtype=torch.int64)
        self.temperature = mk(R, dtype=torch.float32)
        self.seeds = mk(R, dtype=torch.int64)
        self.steps = mk(R, dtype=torch.int32)
        self.fix_idx = mk(R, dtype=torch.int64)
        self.fix_rows = mk(R, dtype=torch.int64)
        self.rows = mk(R, dtype=torch.int64)
        # GPU-side input prep: [rows | num_computed | num_scheduled | append_pos] packed for ONE H2D copy per step, and
        # the staged block-table appends [row | start | count | offset] + their ids
        self.meta = mk(4 * R, dtype=torch.int32)
        self.tok_hdr = mk(4 * R, dtype=torch.int32)
        self.diff_hdr = [mk(8 * R, dtype=torch.int32) for _ in range(G)] if pin else []
        self.diff_ids = [mk(min(R * MB, max(1 << 20, 4 * R + MB)), dtype=torch.int32) for _ in range(G)] if pin else []
        self.event = torch.cuda.Event() if pin else None

    def wait(self) -> None:
        if self.event is not None:
            self.event.synchronize()


class AsyncStepOutput:
    """Result of a launched step; ``get()`` waits for the D2H copy and does the host-side bookkeeping."""

    def __init__(self, runner: "ModelRunner", req_ids: list[str], produces: np.ndarray, staging: _Staging,
                 nb: int, event, t0: float, lp_out=None, plp=None) -> None:
        self.runner, self.req_ids, self.produces = runner, req_ids, produces
        self.staging, self.nb, self.event, self.t0 = staging, nb, event, t0
        self.lp_out, self.plp = lp_out, plp or {}
        self._out: Optional[ModelRunnerOutput] = None

    def done(self) -> bool:
        return self._out is not None or self.event is None or self.event.query()

    def get(self) -> ModelRunnerOutput:
        if self._out is not None:
            return self._out
        r = self.runner
        if self.event is not None:
            t_w = time.perf_counter()
            self.event.synchronize()
            r.wait_s += time.perf_counter() - t_w
        toks = self.staging.sampled[:self.nb].tolist()
        r.d2h_bytes += self.nb * 8
        out_ids: list[list[int]] = []
        for j, rid in enumerate(self.req_ids):
            if not self.produces[j]:
                out_ids.append([])
                continue
            t = toks[j]
            out_ids.append([t])
            row = r.req_index.get(rid)
            if row is None:  # aborted / finished while the step was in flight
                continue
            n = r.num_tokens_cpu[row]
            r.token_ids_cpu[row, n] = t
            r.num_tokens_cpu[row] = n + 1
            r.states[rid].output_token_ids.append(t)
        spec_ids = None
        if r.drafter is not None:
            rows = [r.req_index[rid] for rid in self.req_ids if rid in r.req_index]
            if len(rows) == len(self.req_ids):
                spec_ids = r.drafter.propose(r, rows, out_ids)
        if r._pending is self:
            r._pending = None
        self._out = ModelRunnerOutput(
            req_ids=self.req_ids, req_id_to_index={x: i for i, x in enumerate(self.req_ids)},
            sampled_token_ids=out_ids, logprobs=self.lp_out, prompt_logprobs_dict=self.plp,
            spec_token_ids=spec_ids, forward_ms=(time.perf_counter() - self.t0) * 1e3,
            kv_connector_output=r.kv_connector.get_finished() if r.kv_connector is not None else None)
        return self._out


class ModelRunner:
    def __init__(self, vllm_config: VllmConfig, device: torch.device) -> None:
        self.vllm_config = vllm_config
        self.mc = vllm_config.model_config
        self.sc = vllm_config.scheduler_config
        self.cc = vllm_config.cache_config
        self.device = device
        self.is_cuda = device.type == "cuda"
        self.dtype = self.mc.torch_dtype
        self.max_num_reqs = self.sc.max_num_seqs
        self.max_num_tokens = self.sc.max_num_batched_tokens
        self.max_model_len = self.sc.max_model_len
        self.block_size = self.cc.block_size
        # decode context parallelism: one block-table entry covers block_size * dcp sequence positions, of which
        # this rank stores every dcp-th one (layers/attention.py MLAAttentionCore._forward_dcp)
        self.dcp = max(vllm_config.parallel_config.decode_context_parallel_size, 1)
        self.dcp_rank = 0
        # prefill context parallelism: every rank of the PCP group runs the per-token layers on 1/P of the step's tokens
        # (head + tail chunk of each long prefill for causal balance), K/V are all-gathered per layer so every rank's
        # cache is complete (reference: v1/worker/gpu/pcp_manager.py DualChunkSwap, "partial query / full KV")
        self.pcp = max(vllm_config.parallel_config.prefill_context_parallel_size, 1)
        if self.pcp > 1 and vllm_config.parallel_config.enable_expert_parallel:
            raise NotImplementedError("prefill context parallelism with --enable-expert-parallel")
        self.vblock = self.block_size * self.dcp
        self.max_blocks_per_req = cdiv(self.max_model_len, self.vblock) + 1
        self.vocab_size = self.mc.vocab_size
        self.model: Optional[torch.nn.Module] = None
        from vllm_b200.sample.sampler import load_logits_processors
        self.sampler = Sampler(load_logits_processors(vllm_config.model_config.logits_processors, vllm_config, device))
        self.rejection_sampler = RejectionSampler(self.sampler)
        self.kv_caches: dict[str, torch.Tensor] = {}
        self.layer_to_group: dict[str, int] = {}
        self.num_kv_groups = 1
        self.kv_cache_config: Optional[KVCacheConfig] = None
        self.kv_vt: list[bool] = [False]
        spec = vllm_config.speculative_config
        self.num_spec_tokens = spec.num_speculative_tokens if spec else 0
        self.drafter = None
        self.spec_hidden = None
        self.step_hidden = None
        self._last_aux = None
        import os as _os
        self.cascade_min_tokens = int(_os.environ.get("VLLM_B200_CASCADE_MIN_TOKENS", "256"))
        self.cascade_min_reqs = int(_os.environ.get("VLLM_B200_CASCADE_MIN_REQS", "8"))
        pc_ = vllm_config.parallel_config
        self.eplb_interval = pc_.eplb_step_interval if pc_.enable_eplb else 0
        self.eplb_hook = None             # set by the worker (needs the process groups)
        self.cross_states: dict[str, Any] = {}       # encoder-decoder: request id → per-layer cross-attention (K, V)
        self.is_enc_dec = bool(vllm_config.model_config.hf_config.get("is_encoder_decoder"))
        self.encoder_cache = None
        self.ec_connector = None
        ecp = os.environ.get("VLLM_B200_EC_SHARED_PATH")
        if ecp:
            from vllm_b200.ec_transfer import ECSharedStorageConnector
            self.ec_connector = ECSharedStorageConnector(ecp, os.environ.get("VLLM_B200_EC_ROLE", "ec_both"))
        from vllm_b200.multimodal import MULTIMODAL_REGISTRY
        self.has_mm = MULTIMODAL_REGISTRY.supports_multimodal(vllm_config.model_config)
        self.uses_mrope = vllm_config.model_config.hf_config.get("model_type") in ("qwen2_vl", "qwen2_5_vl", "qwen3_vl", "qwen3_vl_moe", "qwen3_5", "qwen3_5_moe")
        self._deepstack, self._deepstack_live = None, False
        self.kv_connector = None
        ktc = vllm_config.kv_transfer_config
        if ktc is not None and ktc.kv_connector:
            from vllm_b200.kv_transfer import create_connector
            self.kv_connector = create_connector(vllm_config, role="worker")
        self.async_mode = bool(self.sc.async_scheduling)
        # ---- persistent batch (CPU)
        R, L = self.max_num_reqs, self.max_model_len
        self.req_ids: list[str] = []
        self.req_index: dict[str, int] = {}
        self.states: dict[str, ReqState] = {}
        self._parked: dict[str, ReqState] = {}
        self.token_ids_cpu = np.zeros((R, L + self.num_spec_tokens + 1), dtype=np.int32)
        self.num_tokens_cpu = np.zeros(R, dtype=np.int32)       # tokens materialised on the host
        self.num_logical_cpu = np.zeros(R, dtype=np.int32)      # incl. tokens still in flight (async)
        self.num_prompt_cpu = np.zeros(R, dtype=np.int32)
        self.num_computed_cpu = np.zeros(R, dtype=np.int32)
        self.temp_cpu = np.zeros(R, dtype=np.float32)
        self.seed_cpu = np.zeros(R, dtype=np.int64)
        self.complex_cpu = np.zeros(R, dtype=bool)
        self.filter_cpu = np.zeros(R, dtype=bool)                # top-k / top-p / min-p
        self.block_tables_cpu: list[np.ndarray] = []
        self.num_blocks_cpu: list[np.ndarray] = []
        # ---- device buffers
        T = self.max_num_tokens
        dev = self.device
        self.input_ids = torch.zeros(T, dtype=torch.int32, device=dev)
        self.positions = torch.zeros(T, dtype=torch.int64, device=dev)
        self.qsl = torch.zeros(R + 1, dtype=torch.int32, device=dev)
        self.seq_lens = torch.zeros(R, dtype=torch.int32, device=dev)
        self.last_sampled = torch.zeros(R, dtype=torch.int64, device=dev)
        self.slot_mappings: list[torch.Tensor] = []
        self.block_tables: list[torch.Tensor] = []
        self.staging: list[_Staging] = []
        self._stage_i = 0
        self._pending: Optional[AsyncStepOutput] = None
        self.graphs: dict[int, Any] = {}
        self.mixed_graphs: dict[int, Any] = {}     # token bucket → MixedGraph (cudagraph_mode FULL)
        self.graph_pool = None
        self.use_cuda_graph = (self.is_cuda and not self.mc.enforce_eager
                               and vllm_config.compilation_config.cudagraph_mode != "NONE" and self.pcp == 1)
        self.capture_sizes: list[int] = []
        self._step = 0
        self.launch_count = 0   # our kernels launched (graph replays count their captured launches)
        # device-resident request state + on-device expansion of the step inputs (csrc/input_prep.cu); decided in
        # initialize_kv_cache
        self.gpu_prep = False
        self.dev_tokens: Optional[torch.Tensor] = None
        self.dev_block_rows: list[torch.Tensor] = []
        self._bt_diffs: list[list[tuple[str, int, list[int]]]] = []
        self._new_rids: list[str] = []
        self.prep_host_s = 0.0  # host seconds spent preparing + uploading step inputs (diagnostic)
        self.prep_decode_s, self.prep_decode_n = 0.0, 0
        self.wait_s = 0.0       # host seconds blocked on the device (D2H of sampled ids, staging reuse)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # ------------------------------------------------------------------ setup
    def load_model(self, weights=None) -> None:
        from vllm_b200.model_loader import get_model
        clear_attention_layers()
        self.model = get_model(self.vllm_config, self.device, weights)
        if self.vllm_config.speculative_config is not None:
            from vllm_b200.spec_decode import make_drafter
            self.drafter = make_drafter(self.vllm_config, self)
        self.spec_hidden = None
        if getattr(self.drafter, "needs_hidden", False):
            self.spec_hidden = torch.zeros(self.max_num_reqs, self.vllm_config.model_config.hidden_size,
                                           dtype=self.dtype, device=self.device)
        self.lora_manager = None
        if self.vllm_config.lora_config is not None:
            from vllm_b200.lora import LoRAModelManager
            self.lora_manager = LoRAModelManager(self.model, self.vllm_config.lora_config, self.device, self.dtype)

    def get_kv_cache_specs(self) -> dict[str, KVCacheSpec]:
        dsz = self._kv_dtype().itemsize
        return {name: layer.get_kv_cache_spec(self.block_size, dsz) for name, layer in attention_layers().items()}

    def _kv_dtype(self) -> torch.dtype:
        cd = self.cc.cache_dtype
        if cd == "auto":
            return self.dtype
        if cd.startswith("fp8"):
            return torch.float8_e5m2 if cd.endswith("e5m2") else torch.float8_e4m3fn
        return {"bfloat16": torch.bfloat16, "float16": torch.float16}[cd]

    def initialize_kv_cache(self, cfg: KVCacheConfig) -> None:
        from vllm_b200.core.kv_cache_manager import MambaSpec, MLAAttentionSpec
        self.kv_cache_config = cfg
        if self.dcp > 1:
            from vllm_b200.parallel.state import get_dcp_group
            self.dcp_rank = max(get_dcp_group().rank_in_group, 0)
            hf = self.mc.hf_config
            total_kv = hf.get("num_key_value_heads") or hf.get("num_attention_heads") or 1
            tp = self.vllm_config.parallel_config.tensor_parallel_size
            if any(not isinstance(g.spec, MLAAttentionSpec) for g in cfg.kv_cache_groups) and \
                    (tp % total_kv != 0 or (tp // total_kv) % self.dcp != 0):
                raise ValueError(f"decode context parallelism shards the positions of REPLICATED KV heads: it needs "
                                 f"tensor_parallel_size / num_key_value_heads ({tp} / {total_kv}) to be a multiple of "
                                 f"decode_context_parallel_size ({self.dcp})")
            if not all(isinstance(g.spec, MLAAttentionSpec) or getattr(g.spec, "dcp", 1) == self.dcp
                       for g in cfg.kv_cache_groups):
                raise NotImplementedError("decode context parallelism covers MLA latent caches and full-attention GQA / MQA "
                                          "caches (no sliding-window or state-space groups)")
            if self.use_cuda_graph:
                logger.warning("decode context parallelism runs eagerly (per-request merge on the host)")
                self.use_cuda_graph = False
        self.num_kv_groups = max(len(cfg.kv_cache_groups), 1)
        kvd = self._kv_dtype()
        from vllm_b200 import envs
        self.kv_vt = []
        for gid, g in enumerate(cfg.kv_cache_groups):
            s = g.spec
            # tensor-core cache layout (V pages transposed) whenever the tcgen05 attention kernel applies
            self.kv_vt.append(bool(
                self.is_cuda and not envs.VLLM_B200_FORCE_REF_OPS and envs.VLLM_B200_ATTN != "simt"
                and not isinstance(s, MLAAttentionSpec) and getattr(s, "head_size", 0) == 128
                and s.block_size == 64 and getattr(s, "dcp", 1) == 1 and kvd in (torch.bfloat16, torch.float16)
                and self.dtype in (torch.bfloat16, torch.float16)
                # features only the generic kernel implements keep the whole group on the standard layout
                and all(not getattr(attention_layers().get(n), "logits_soft_cap", 0.0)
                        and getattr(attention_layers().get(n), "alibi_slopes", None) is None
                        and getattr(attention_layers().get(n), "sinks", None) is None
                        and not getattr(attention_layers().get(n), "chunk_size", 0) for n in g.layer_names)))
            for name in g.layer_names:
                self.layer_to_group[name] = gid
                if isinstance(s, MambaSpec):
                    self.kv_caches[name] = torch.zeros((cfg.num_blocks, s.state_bytes // 4), dtype=torch.float32,
                                                       device=self.device)
                    continue
                if isinstance(s, MLAAttentionSpec):
                    shape = (cfg.num_blocks, s.phys_block_size, s.head_size)
                else:
                    shape = (cfg.num_blocks, 2, s.num_kv_heads, getattr(s, "phys_block_size", s.block_size), s.head_size)
                self.kv_caches[name] = torch.zeros(shape, dtype=kvd, device=self.device)
        if self.kv_connector is not None and hasattr(self.kv_connector, "register_kv_caches"):
            self.kv_connector.register_kv_caches(self.kv_caches)
        G, R, MB = self.num_kv_groups, self.max_num_reqs, self.max_blocks_per_req
        self.block_tables_cpu = [np.zeros((R, MB), dtype=np.int32) for _ in range(G)]
        self.num_blocks_cpu = [np.zeros(R, dtype=np.int32) for _ in range(G)]
        self.block_tables = [torch.zeros(R, MB, dtype=torch.int32, device=self.device) for _ in range(G)]
        self.slot_mappings = [torch.zeros(self.max_num_tokens, dtype=torch.int64, device=self.device)
                              for _ in range(G)]
        n_sets = 3 if self.is_cuda else 1
        self.staging = [_Staging(self.max_num_tokens, R, MB, G, self.is_cuda) for _ in range(n_sets)]
        self._init_gpu_prep()

    def _init_gpu_prep(self) -> None:
        """Persistent device mirrors of the per-request token ids and block tables (reference:
        v1/worker/gpu/states.py:9, block_table.py:228-330).  With them a step uploads four ints per scheduled request
        plus the block ids that were appended since the last step; token ids, positions, slot mappings, query offsets
        and the compact block table are expanded by ``prep_step`` on the device, and sampled ids are appended to the
        table by ``append_sampled`` without a host round trip."""
        from vllm_b200 import envs
        pc = self.vllm_config.parallel_config
        R, L = self.max_num_reqs, self.token_ids_cpu.shape[1]
        ok = (self.is_cuda and envs.VLLM_B200_GPU_INPUT_PREP and not envs.VLLM_B200_FORCE_REF_OPS and self.dcp == 1
              and self.pcp == 1 and not self.uses_mrope and self.num_spec_tokens == 0 and not self.is_enc_dec
              and pc.pipeline_parallel_size == 1 and R * L * 4 <= (1 << 30))
        self.gpu_prep = bool(ok)
        if not ok:
            return
        from vllm_b200.ops import cuda_ops
        self._cu = cuda_ops()
        self.dev_tokens = torch.zeros(R, L, dtype=torch.int32, device=self.device)
        self.dev_block_rows = [torch.zeros(R, self.max_blocks_per_req, dtype=torch.int32, device=self.device)
                               for _ in range(self.num_kv_groups)]
        self.step_meta = torch.zeros(4 * R, dtype=torch.int32, device=self.device)
        self.tok_hdr_dev = torch.zeros(4 * R, dtype=torch.int32, device=self.device)
        self.diff_hdr_dev = [torch.zeros(8 * R, dtype=torch.int32, device=self.device) for _ in range(self.num_kv_groups)]
        self.diff_ids_dev = [torch.zeros(self.staging[0].diff_ids[g].numel(), dtype=torch.int32, device=self.device)
                             for g in range(self.num_kv_groups)]
        self._bt_diffs = [[] for _ in range(self.num_kv_groups)]

    # ------------------------------------------------------- batch maintenance
    def _move_row(self, src: int, dst: int) -> None:
        n = self.num_tokens_cpu[src]
        self.token_ids_cpu[dst, :n] = self.token_ids_cpu[src, :n]
        for a in (self.num_tokens_cpu, self.num_logical_cpu, self.num_prompt_cpu, self.num_computed_cpu,
                  self.temp_cpu, self.seed_cpu, self.complex_cpu, self.filter_cpu):
            a[dst] = a[src]
        for g in range(self.num_kv_groups):
            nb = self.num_blocks_cpu[g][src]
            self.block_tables_cpu[g][dst, :nb] = self.block_tables_cpu[g][src, :nb]
            self.num_blocks_cpu[g][dst] = nb
        if self.is_cuda:
            self.last_sampled[dst:dst + 1].copy_(self.last_sampled[src:src + 1])
        if self.gpu_prep:       # stream-ordered after the append of the step that is still in flight
            nl = int(self.num_logical_cpu[src]) + 1
            self.dev_tokens[dst, :nl].copy_(self.dev_tokens[src, :nl])
            for g in range(self.num_kv_groups):
                nb = int(self.num_blocks_cpu[g][src])
                self.dev_block_rows[g][dst, :nb].copy_(self.dev_block_rows[g][src, :nb])

    def _remove(self, req_id: str, park: bool = False) -> None:
        row = self.req_index.pop(req_id, None)
        st = self.states.pop(req_id, None)
        if park and st is not None:
            self._parked[req_id] = st
        elif self.cross_states:
            self.cross_states.pop(req_id, None)
        if row is None:
            return
        last = len(self.req_ids) - 1
        if row != last:  # move the last row into the hole (keeps rows dense)
            mv = self.req_ids[last]
            self.req_ids[row] = mv
            self.req_index[mv] = row
            self.states[mv].row = row
            self._move_row(last, row)
        self.req_ids.pop()

    def _add_row(self, st: ReqState, token_ids, block_ids, num_computed: int) -> None:
        row = len(self.req_ids)
        self.req_ids.append(st.req_id)
        self.req_index[st.req_id] = row
        st.row = row
        self.states[st.req_id] = st
        n = len(token_ids)
        self.token_ids_cpu[row, :n] = token_ids
        self.num_tokens_cpu[row] = n
        self.num_logical_cpu[row] = n
        self.num_prompt_cpu[row] = st.num_prompt_tokens
        self.num_computed_cpu[row] = num_computed
        sp = st.sampling_params
        if sp is not None:
            self.temp_cpu[row] = sp.temperature
            # unseeded requests get a seed that is identical on every TP rank (crc of the id + engine seed)
            self.seed_cpu[row] = (sp.seed if sp.seed is not None
                                  else zlib.crc32(st.req_id.encode()) * 2654435761 + self.mc.seed) & 0x7FFFFFFFFFFF
            self.filter_cpu[row] = sp.temperature >= 1e-5 and (sp.top_k > 0 or sp.top_p < 1.0 or sp.min_p > 0)
            st.complex_sampling = bool(
                sp.needs_penalties or sp.logprobs is not None or sp.prompt_logprobs is not None or sp.logit_bias
                or sp.allowed_token_ids or sp.bad_words_token_ids or sp.min_tokens > 0
                or sp.structured_outputs is not None
                or (self.sampler.custom and any(k != "kv_transfer_params" for k in (sp.extra_args or {}))))
        else:
            self.temp_cpu[row] = 0.0
            self.seed_cpu[row] = 0
            self.filter_cpu[row] = False
            st.complex_sampling = True
        self.complex_cpu[row] = st.complex_sampling
        for g in range(self.num_kv_groups):
            ids = block_ids[g]
            self.block_tables_cpu[g][row, :len(ids)] = ids
            self.num_blocks_cpu[g][row] = len(ids)
        if self.gpu_prep:       # admission: the whole row travels once, with this step's uploads (_flush_gpu_state)
            self._new_rids.append(st.req_id)

    def _add(self, nr) -> None:
        sp = nr.sampling_params
        gen = None
        if sp is not None and sp.seed is not None:
            gen = torch.Generator(device=self.device)
            gen.manual_seed(sp.seed)
        st = ReqState(nr.req_id, -1, sp, nr.pooling_params, gen, nr.lora_request, len(nr.prompt_token_ids),
                      mm_features=nr.mm_features)
        self._add_row(st, nr.prompt_token_ids, nr.block_ids, nr.num_computed_tokens)

    def _update_states(self, so: SchedulerOutput) -> None:
        for rid in so.finished_req_ids:
            self._remove(rid)
            self._parked.pop(rid, None)
        for rid in so.preempted_req_ids:
            self._remove(rid, park=True)
        for nr in so.scheduled_new_reqs:
            if nr.req_id in self.req_index:
                self._remove(nr.req_id)
            self._add(nr)
        c = so.scheduled_cached_reqs
        for i, rid in enumerate(c.req_ids):
            if c.resumed_from_preemption[i] or rid not in self.req_index:
                # re-admit with its full token history (prompt + outputs so far)
                st = self._parked.pop(rid, None) or self.states.get(rid)
                assert st is not None, f"resumed request {rid} has no parked state"
                self._remove(rid)
                self._add_row(st, c.new_token_ids[i], c.new_block_ids[i], c.num_computed_tokens[i])
                continue
            row = self.req_index[rid]
            self.num_computed_cpu[row] = c.num_computed_tokens[i]
            toks = c.new_token_ids[i]
            if toks:  # PP: tokens sampled on another rank
                n0 = c.num_computed_tokens[i]
                self.token_ids_cpu[row, n0:n0 + len(toks)] = toks
                self.num_tokens_cpu[row] = max(self.num_tokens_cpu[row], n0 + len(toks))
                self.num_logical_cpu[row] = max(self.num_logical_cpu[row], self.num_tokens_cpu[row])
            nb = c.new_block_ids[i]
            if nb is not None:
                for g in range(self.num_kv_groups):
                    ids = nb[g]
                    if ids:
                        s = self.num_blocks_cpu[g][row]
                        self.block_tables_cpu[g][row, s:s + len(ids)] = ids
                        self.num_blocks_cpu[g][row] = s + len(ids)
                        if self.gpu_prep:     # keyed by request: a later removal in this step may still move the row
                            self._bt_diffs[g].append((rid, int(s), ids))

    # ---------------------------------------------------------- input prep
    def _prepare(self, so: SchedulerOutput, stg: _Staging):
        B = len(self.req_ids)
        get = so.num_scheduled_tokens.get
        n_sched = np.fromiter((get(r, 0) for r in self.req_ids), dtype=np.int32, count=B)
        rows = np.nonzero(n_sched)[0].astype(np.int64)
        ns = n_sched[rows]
        nb = len(rows)
        T = int(ns.sum())
        cu = np.zeros(nb + 1, dtype=np.int32)
        np.cumsum(ns, out=cu[1:])
        computed = self.num_computed_cpu[rows].astype(np.int64)
        if T == nb:  # pure decode: one token per row
            req_of_tok = np.arange(nb)
            pos = computed
        else:
            req_of_tok = np.repeat(np.arange(nb), ns)
            offs = np.arange(T, dtype=np.int64) - np.repeat(cu[:-1].astype(np.int64), ns)
            pos = computed[req_of_tok] + offs
        row_of_tok = rows[req_of_tok]
        # speculative tokens are appended after the known tokens of the row
        for rid, drafts in so.scheduled_spec_decode_tokens.items():
            r = self.req_index[rid]
            n = self.num_tokens_cpu[r]
            self.token_ids_cpu[r, n:n + len(drafts)] = drafts
        ids = self.token_ids_cpu[row_of_tok, pos]
        # tokens sampled by a step that is still in flight: patched on the device from last_sampled
        n_spec_row = np.zeros(nb, dtype=np.int64)
        unknown = pos >= (self.num_tokens_cpu[row_of_tok].astype(np.int64))
        if so.scheduled_spec_decode_tokens:
            unknown[:] = False
        fix = np.nonzero(unknown)[0]
        stg.input_ids[:T] = torch.from_numpy(ids)
        rope_pos = pos
        if self.uses_mrope:
            # M-RoPE: image placeholders compress many tokens into few position steps, so a request's rotary position
            # runs behind its sequence index by a per-request delta once the prompt is past
            delta = np.array([self._mrope_delta(self.req_ids[r]) for r in rows], dtype=np.int64)
            rope_pos = pos + delta[req_of_tok]
        stg.positions[:T] = torch.from_numpy(rope_pos)
        seq_lens = (computed + ns).astype(np.int32)
        stg.qsl[:nb + 1] = torch.from_numpy(cu)
        stg.seq_lens[:nb] = torch.from_numpy(seq_lens)
        stg.rows[:nb] = torch.from_numpy(rows)
        bs = self.block_size
        for g in range(self.num_kv_groups):
            bt = self.block_tables_cpu[g]
            if self.dcp > 1:
                d, vb = self.dcp, self.vblock
                slots = bt[row_of_tok, pos // vb].astype(np.int64) * bs + (pos % vb) // d
                slots = np.where(pos % d == self.dcp_rank, slots, -1)      # positions stored by another rank
            else:
                slots = bt[row_of_tok, pos // bs].astype(np.int64) * bs + pos % bs
            stg.slots[g][:T] = torch.from_numpy(slots)
        if len(fix):
            stg.fix_idx[:len(fix)] = torch.from_numpy(fix)
            stg.fix_rows[:len(fix)] = torch.from_numpy(row_of_tok[fix])
        del n_spec_row
        max_q = int(ns.max()) if nb else 0
        max_s = int(seq_lens.max()) if nb else 0
        return rows, ns, cu, T, max_q, max_s, len(fix)

    def _h2d(self, stg: _Staging, rows, nb, T, pad_b, pad_t, max_s, n_fix) -> None:
        nbk = cdiv(max_s, self.vblock)
        self.h2d_bytes += T * (4 + 8) + (nb + 1) * 4 + nb * 4 + self.num_kv_groups * (T * 8 + nb * nbk * 4)
        nbc = True
        self.input_ids[:T].copy_(stg.input_ids[:T], non_blocking=nbc)
        self.positions[:T].copy_(stg.positions[:T], non_blocking=nbc)
        self.qsl[:nb + 1].copy_(stg.qsl[:nb + 1], non_blocking=nbc)
        self.seq_lens[:nb].copy_(stg.seq_lens[:nb], non_blocking=nbc)
        for g in range(self.num_kv_groups):
            self.slot_mappings[g][:T].copy_(stg.slots[g][:T], non_blocking=nbc)
            stg.block_tables[g][:nb, :nbk] = torch.from_numpy(self.block_tables_cpu[g][rows, :nbk])
            self.block_tables[g][:nb, :nbk].copy_(stg.block_tables[g][:nb, :nbk], non_blocking=nbc)
        if n_fix:
            idx = stg.fix_idx[:n_fix].to(self.device, non_blocking=True)
            src = stg.fix_rows[:n_fix].to(self.device, non_blocking=True)
            self.input_ids[:T].index_copy_(0, idx, self.last_sampled[src].to(torch.int32))
            self.h2d_bytes += n_fix * 16
        if pad_t > T:  # padded decode slots: write to the null block, attend over 1 token of block 0
            self.input_ids[T:pad_t].zero_()
            self.positions[T:pad_t].zero_()
            for g in range(self.num_kv_groups):
                self.slot_mappings[g][T:pad_t].fill_(-1)
                self.block_tables[g][nb:pad_b, :1].zero_()
            self.seq_lens[nb:pad_b].fill_(1)
            self.qsl[nb + 1:pad_b + 1].copy_(
                torch.arange(T + 1, T + 1 + pad_b - nb, dtype=torch.int32, device=self.device))

    # ------------------------------------------------- device-side input prep
    def _flush_gpu_state(self, stg: _Staging) -> None:
        """Bring the device mirrors up to date with O(1) CUDA calls: the token rows of requests admitted this step and
        the block ids appended since the last step are packed into pinned staging, shipped with one copy each, and
        scattered into the tables by ``block_table_apply_diffs`` (row, start, count, offset) entries."""
        cu = self._cu
        R = self.max_num_reqs
        if self._new_rids:
            rids, self._new_rids = self._new_rids, []
            pin, cap = stg.input_ids, stg.input_ids.numel()
            pv = pin.numpy()
            d_row, d_cnt, d_off = [], [], []
            off = 0
            for rid in rids:
                row = self.req_index.get(rid)
                if row is None:
                    continue
                n = int(self.num_tokens_cpu[row])
                for g in range(self.num_kv_groups):      # the whole block list of the row goes through the diff path
                    k = int(self.num_blocks_cpu[g][row])
                    if k:
                        self._bt_diffs[g].append((rid, 0, self.block_tables_cpu[g][row, :k]))
                if off + n <= cap and len(d_row) < R:
                    pv[off:off + n] = self.token_ids_cpu[row, :n]
                    d_row.append(row)
                    d_cnt.append(n)
                    d_off.append(off)
                    off += n
                else:                       # does not fit the packed upload (very long prompts): pageable row copy
                    self.dev_tokens[row, :n].copy_(torch.from_numpy(self.token_ids_cpu[row, :n].copy()))
                self.h2d_bytes += n * 4
            n_e = len(d_row)
            if n_e:
                hv = stg.tok_hdr.numpy()
                hv[:n_e], hv[n_e:2 * n_e], hv[2 * n_e:3 * n_e], hv[3 * n_e:4 * n_e] = d_row, 0, d_cnt, d_off
                self.tok_hdr_dev[:4 * n_e].copy_(stg.tok_hdr[:4 * n_e], non_blocking=True)
                # the step's own input_ids buffer doubles as device scratch: prep_step overwrites it afterwards
                self.input_ids[:off].copy_(pin[:off], non_blocking=True)
                th = self.tok_hdr_dev
                cu.block_table_apply_diffs(self.dev_tokens, th[:n_e], th[n_e:2 * n_e], th[2 * n_e:3 * n_e],
                                           th[3 * n_e:4 * n_e], self.input_ids)
        for g in range(self.num_kv_groups):
            diffs = self._bt_diffs[g]
            if not diffs:
                continue
            iv = stg.diff_ids[g].numpy()
            d_row, d_start, d_cnt, d_off = [], [], [], []
            o = 0
            for rid, start, new in diffs:
                row = self.req_index.get(rid)
                if row is None:
                    continue
                k = len(new)
                if len(d_row) >= 2 * R or o + k > iv.shape[0]:   # staging full (pathological): upload the row
                    tot = int(self.num_blocks_cpu[g][row])
                    self.dev_block_rows[g][row, :tot].copy_(torch.from_numpy(self.block_tables_cpu[g][row, :tot].copy()))
                    continue
                d_row.append(row)
                d_start.append(start)
                d_cnt.append(k)
                d_off.append(o)
                iv[o:o + k] = new
                o += k
            diffs.clear()
            n = len(d_row)
            if n == 0:
                continue
            hv = stg.diff_hdr[g].numpy()
            hv[:n], hv[n:2 * n], hv[2 * n:3 * n], hv[3 * n:4 * n] = d_row, d_start, d_cnt, d_off
            dh, di = self.diff_hdr_dev[g], self.diff_ids_dev[g]
            dh[:4 * n].copy_(stg.diff_hdr[g][:4 * n], non_blocking=True)
            di[:o].copy_(stg.diff_ids[g][:o], non_blocking=True)
            self.h2d_bytes += 16 * n + 4 * o
            cu.block_table_apply_diffs(self.dev_block_rows[g], dh[:n], dh[n:2 * n], dh[2 * n:3 * n], dh[3 * n:4 * n], di)

    def _prepare_gpu(self, so: SchedulerOutput, stg: _Staging):
        """O(requests) host work: which rows run, how many tokens each, and the scalars the device expansion needs."""
        B = len(self.req_ids)
        get = so.num_scheduled_tokens.get
        n_sched = np.fromiter((get(r, 0) for r in self.req_ids), dtype=np.int32, count=B)
        rows = np.nonzero(n_sched)[0].astype(np.int64)
        ns = n_sched[rows]
        nb = len(rows)
        T = int(ns.sum())
        cu = np.zeros(nb + 1, dtype=np.int32)
        np.cumsum(ns, out=cu[1:])
        computed = self.num_computed_cpu[rows]
        seq_lens = computed + ns
        produces = seq_lens >= self.num_logical_cpu[rows]
        m = stg.meta.numpy()
        m[:nb] = rows
        m[nb:2 * nb] = computed
        m[2 * nb:3 * nb] = ns
        m[3 * nb:4 * nb] = np.where(produces, seq_lens, -1)     # where the sampled id goes in the device token table
        stg.rows[:nb] = torch.from_numpy(rows)
        max_q = int(ns.max()) if nb else 0
        max_s = int(seq_lens.max()) if nb else 0
        return rows, ns, cu, T, max_q, max_s, 0

    def _h2d_gpu(self, stg: _Staging, nb, T, pad_b, max_s, pad_t: int = 0) -> None:
        cu = self._cu
        self._flush_gpu_state(stg)
        self.step_meta[:4 * nb].copy_(stg.meta[:4 * nb], non_blocking=True)
        self.h2d_bytes += 16 * nb
        nbk = max(cdiv(max_s, self.vblock), 1)
        sm = self.step_meta
        for g in range(self.num_kv_groups):
            cu.prep_step(sm[:nb], sm[nb:2 * nb], sm[2 * nb:3 * nb], self.dev_tokens, self.dev_block_rows[g], self.input_ids,
                         self.positions, self.qsl, self.seq_lens, self.slot_mappings[g], self.block_tables[g], nbk,
                         self.block_size, pad_b=pad_b, write_tokens=(g == 0), pad_t=pad_t)

    def _make_metadata(self, nb, T, max_q, max_s, decode_only, cpu_qsl=None, cpu_sl=None) -> dict[int, AttentionMetadata]:
        nbk = max(cdiv(max_s, self.vblock), 1)
        out = {}
        for g in range(self.num_kv_groups):
            out[g] = AttentionMetadata(
                num_reqs=nb, num_actual_tokens=T, max_query_len=max_q, max_seq_len=max_s,
                query_start_loc=self.qsl[:nb + 1], seq_lens=self.seq_lens[:nb],
                block_table=self.block_tables[g][:nb] if self.is_cuda else self.block_tables[g][:nb, :nbk],
                slot_mapping=self.slot_mappings[g][:T], query_start_loc_cpu=cpu_qsl, seq_lens_cpu=cpu_sl,
                decode_only=decode_only, kv_vt=self.kv_vt[g])
        return out

    # ------------------------------------------------------------ forward
    def _cross_context(self, states, cu):
        """Encoder-decoder models: run the encoder for requests seen for the first time and hand the decoder the
        per-request cross-attention K/V (kept on the device for the life of the request)."""
        kvs = []
        for s in states:
            kv = self.cross_states.get(s.req_id)
            if kv is None and s.mm_features:
                f = next((f for f in s.mm_features if f.get("modality") in ("encoder_tokens", "audio")), None)
                if f is not None:
                    src = (torch.as_tensor(f["data"]).to(self.device) if f["modality"] == "audio"
                           else torch.tensor(f["data"], dtype=torch.int64, device=self.device))
                    kv = self.model.encode(src)
                    self.cross_states[s.req_id] = kv
            kvs.append(kv)
        return [int(x) for x in cu], kvs

    def _forward(self, T: int, md: dict[int, AttentionMetadata], intermediate=None, inputs_embeds=None, extra=None):
        ctx = ForwardContext(attn_metadata=md, kv_caches=self.kv_caches, layer_to_group=self.layer_to_group,
                             num_tokens=T)
        if extra:
            ctx.extra.update(extra)
        with set_forward_context(ctx):
            if inputs_embeds is not None:
                out = self.model(self.input_ids[:T], self.positions[:T], intermediate, inputs_embeds)
            else:
                out = self.model(self.input_ids[:T], self.positions[:T], intermediate)
        self._last_aux = ctx.extra.get("aux_hidden")            # EAGLE-3 taps (None otherwise)
        return out

    # ------------------------------------------------------------ cascade attention
    def _add_cascade(self, md, common_blocks, rows, ns, nb, T, max_q, max_s) -> None:
        from vllm_b200.core.kv_cache_manager import FullAttentionSpec, SlidingWindowSpec
        computed = self.num_computed_cpu[rows]
        for g in range(self.num_kv_groups):
            spec = self.kv_cache_config.kv_cache_groups[g].spec
            if type(spec) is not FullAttentionSpec or isinstance(spec, SlidingWindowSpec):
                continue
            nblk = int(common_blocks[g]) if g < len(common_blocks) else 0
            P = nblk * self.block_size
            # every query of the step must sit behind the shared prefix, and the prefix must be worth a second pass
            if P < self.cascade_min_tokens or int(computed.min()) < P:
                continue
            m = md[g]
            dev = self.device
            qsl_p = torch.tensor([0, T], dtype=torch.int32)
            sl_p = torch.tensor([P], dtype=torch.int32)
            md_p = AttentionMetadata(
                num_reqs=1, num_actual_tokens=T, max_query_len=T, max_seq_len=P, query_start_loc=qsl_p.to(dev),
                seq_lens=sl_p.to(dev), block_table=m.block_table[:1, :nblk], slot_mapping=m.slot_mapping,
                query_start_loc_cpu=None if self.is_cuda else qsl_p, seq_lens_cpu=None if self.is_cuda else sl_p,
                decode_only=False, kv_vt=m.kv_vt)
            sl_cpu = (computed + ns - P).astype(np.int32)
            sl_s = torch.from_numpy(sl_cpu)
            md_s = AttentionMetadata(
                num_reqs=nb, num_actual_tokens=T, max_query_len=max_q, max_seq_len=max_s - P,
                query_start_loc=m.query_start_loc, seq_lens=sl_s.to(dev), block_table=m.block_table[:, nblk:],
                slot_mapping=m.slot_mapping, query_start_loc_cpu=m.query_start_loc_cpu,
                seq_lens_cpu=None if self.is_cuda else sl_s, decode_only=m.decode_only, kv_vt=m.kv_vt)
            m.cascade = (md_p, md_s)

    # ------------------------------------------------------------ multimodal rotary positions
    def _mrope_delta(self, req_id: str) -> int:
        f = self._mrope_feature(self.states[req_id])
        return int(f["delta"]) if f is not None else 0

    @staticmethod
    def _mrope_feature(st: "ReqState"):
        for f in st.mm_features or ():
            if f.get("modality") == "mrope":
                return f
        return None

    def _mrope_positions(self, states, rows, ns, cu, T):
        """[3, T] ids for a step that touches a stored multimodal prompt range, else None (1-D positions suffice)."""
        need = False
        for j, s in enumerate(states):
            f = self._mrope_feature(s)
            if f is not None and int(self.num_computed_cpu[rows[j]]) < f["data"].shape[1]:
                need = True
        if not need:
            return None
        out = torch.zeros(3, T, dtype=torch.long)
        for j, s in enumerate(states):
            start, n = int(self.num_computed_cpu[rows[j]]), int(ns[j])
            f = self._mrope_feature(s)
            idx = torch.arange(start, start + n)
            if f is None:
                col = idx[None].expand(3, -1)
            else:
                stored, L = torch.as_tensor(f["data"]), f["data"].shape[1]
                col = torch.where(idx[None] < L, stored[:, idx.clamp(max=L - 1)], idx[None] + int(f["delta"]))
            out[:, int(cu[j]):int(cu[j]) + n] = col
        return out.to(self.device)

    # ------------------------------------------------------------ prefill context parallelism
    @staticmethod
    def pcp_plan(cu: np.ndarray, ns: np.ndarray, P: int):
        """Token → rank assignment for one step.  A request with >= 2P scheduled tokens is cut into 2P chunks and rank r
        takes chunks r and 2P-1-r (causal attention cost grows with position: head + tail balances it); shorter
        requests (decodes) go whole to rank ``i % P``.  Returns per rank a list of (request index, a, b) pieces in
        ascending token order, half-open and relative to the request's scheduled tokens."""
        pieces: list[list[tuple[int, int, int]]] = [[] for _ in range(P)]
        for i, n in enumerate(ns.tolist()):
            if n >= 2 * P:
                c = -(-n // (2 * P))
                for j in range(2 * P):
                    a, b = j * c, min((j + 1) * c, n)
                    if a < b:
                        pieces[j if j < P else 2 * P - 1 - j].append((i, a, b))
            elif n > 0:
                pieces[i % P].append((i, 0, n))
        for r in range(P):
            pieces[r].sort(key=lambda t: (t[0], t[1]))
        return pieces

    def _forward_pcp(self, T: int, nb: int, cu: np.ndarray, ns: np.ndarray, seq_lens: np.ndarray, intermediate=None):
        """Runs the model on this rank's token share; returns the hidden states of every request's LAST scheduled
        token, [nb, H] in request order, identical on every rank of the PCP group."""
        from vllm_b200.parallel.state import get_pcp_group
        grp = get_pcp_group()
        P, r = self.pcp, max(grp.rank_in_group, 0)
        pieces = self.pcp_plan(cu, ns, P)
        sel = [np.concatenate([np.arange(cu[i] + a, cu[i] + b) for i, a, b in pc]) if pc else np.zeros(0, dtype=np.int64)
               for pc in pieces]
        sizes = [int(x.shape[0]) for x in sel]
        order = np.concatenate(sel).astype(np.int64)
        inv = torch.from_numpy(np.argsort(order, kind="stable")).to(self.device)
        mine = pieces[r]
        my_sel = torch.from_numpy(sel[r].astype(np.int64)).to(self.device)
        Tl = sizes[r]
        idle = Tl == 0
        if idle:
            # nothing for this rank in this step (e.g. a lone decode): run one throw-away token so that every kernel and
            # every collective still sees a non-empty tensor; it contributes no rows to the gathers and writes no KV
            my_sel = torch.zeros(1, dtype=torch.long, device=self.device)
            Tl = 1
        # metadata of my pieces: each is a pseudo-request "last (b-a) tokens of a sequence of computed+b tokens"
        computed = seq_lens - ns
        qlens = np.array([b - a for _, a, b in mine], dtype=np.int32)
        sl_loc = np.array([computed[i] + b for i, _, b in mine], dtype=np.int32)
        req_idx = torch.tensor([i for i, _, _ in mine], dtype=torch.long, device=self.device)
        qsl_loc = np.zeros(len(mine) + 1, dtype=np.int32)
        np.cumsum(qlens, out=qsl_loc[1:])
        max_q = int(qlens.max()) if len(mine) else 0
        max_s = int(sl_loc.max()) if len(mine) else 1
        nbk = max(cdiv(max_s, self.vblock), 1)
        qsl_t, sl_t = torch.from_numpy(qsl_loc), torch.from_numpy(sl_loc)
        md = {}
        if idle:
            qsl_t, sl_t, max_q, max_s, nbk = torch.tensor([0, 1], dtype=torch.int32), torch.ones(1, dtype=torch.int32), 1, 1, 1
        for g in range(self.num_kv_groups):
            bt = (torch.zeros_like(self.block_tables[g][:1]) if idle            # null block
                  else self.block_tables[g][:nb].index_select(0, req_idx))
            slots = (torch.full((1,), -1, dtype=self.slot_mappings[g].dtype, device=self.device) if idle
                     else self.slot_mappings[g][:T].index_select(0, my_sel))
            md[g] = AttentionMetadata(
                num_reqs=max(len(mine), 1), num_actual_tokens=Tl, max_query_len=max_q, max_seq_len=max_s,
                query_start_loc=qsl_t.to(self.device), seq_lens=sl_t.to(self.device),
                block_table=bt if self.is_cuda else bt[:, :nbk], slot_mapping=slots,
                query_start_loc_cpu=None if self.is_cuda else qsl_t, seq_lens_cpu=None if self.is_cuda else sl_t,
                decode_only=max_q == 1, kv_vt=self.kv_vt[g])

        def gather(x: torch.Tensor) -> torch.Tensor:
            """[Tl, ...] of every rank → [T, ...] in global token order."""
            mx = max(max(sizes), 1)
            pad = x.new_zeros((mx,) + tuple(x.shape[1:]))
            pad[:sizes[r]] = x[:sizes[r]]
            outs = [torch.empty_like(pad) for _ in range(P)]
            torch.distributed.all_gather(outs, pad, group=grp.device_group)
            return torch.cat([o[:n] for o, n in zip(outs, sizes)], dim=0).index_select(0, inv)

        ctx = ForwardContext(attn_metadata=md, kv_caches=self.kv_caches, layer_to_group=self.layer_to_group, num_tokens=Tl)
        ctx.extra["pcp"] = {"gather": gather, "slot_mappings": [self.slot_mappings[g][:T] for g in range(self.num_kv_groups)],
                            "num_tokens": T}
        with set_forward_context(ctx):
            hidden = self.model(self.input_ids[:T].index_select(0, my_sel), self.positions[:T].index_select(0, my_sel),
                                intermediate)
        if isinstance(hidden, dict):
            return hidden
        # last scheduled token of every request: owned by exactly one rank → exchange just those rows
        last = cu[1:].astype(np.int64) - 1
        owner_pos = {int(t): (rr, k) for rr in range(P) for k, t in enumerate(sel[rr].tolist())}
        mine_last = [owner_pos[int(t)][1] for t in last if owner_pos[int(t)][0] == r]
        cnt = [sum(1 for t in last if owner_pos[int(t)][0] == rr) for rr in range(P)]
        loc = hidden.index_select(0, torch.tensor(mine_last, dtype=torch.long, device=self.device))
        mx = max(max(cnt), 1)
        pad = hidden.new_zeros(mx, hidden.shape[-1])
        pad[:loc.shape[0]] = loc
        outs = [torch.empty_like(pad) for _ in range(P)]
        torch.distributed.all_gather(outs, pad, group=grp.device_group)
        rows_by_rank = [[i for i, t in enumerate(last) if owner_pos[int(t)][0] == rr] for rr in range(P)]
        out = hidden.new_zeros(nb, hidden.shape[-1])
        for rr in range(P):
            if rows_by_rank[rr]:
                out[torch.tensor(rows_by_rank[rr], dtype=torch.long, device=self.device)] = outs[rr][:cnt[rr]]
        return out

    def _mm_inputs_embeds(self, states, rows, ns, cu, T):
        """Token embeddings of the step with multimodal placeholder positions replaced by encoder outputs.
        Returns None when no scheduled token range touches a multimodal item."""
        todo = []
        for j, s in enumerate(states):
            if not s.mm_features:
                continue
            start = int(self.num_computed_cpu[rows[j]])
            end = start + int(ns[j])
            for f in s.mm_features:
                lo, hi = max(start, f["offset"]), min(end, f["offset"] + f["length"])
                if lo < hi:
                    todo.append((j, f, lo, hi, start))
        if not todo:
            return None
        if self.encoder_cache is None:
            from vllm_b200.multimodal import EncoderCache
            self.encoder_cache = EncoderCache()
        emb = self.model.embed_input_ids(self.input_ids[:T]).clone()
        self._deepstack_live = False
        for j, f, lo, hi, start in todo:
            enc = self.encoder_cache.get(f["hash"])
            if enc is None and self.ec_connector is not None:          # another instance may have encoded it
                enc = self.ec_connector.load(f["hash"], self.device)
                if enc is not None:
                    enc = enc.to(emb.dtype)
                    self.encoder_cache.put(f["hash"], enc)
            if enc is None and f.get("modality") == "embeds":            # caller-provided prompt embeddings
                enc = torch.as_tensor(f["data"]).to(self.device, emb.dtype)
                self.encoder_cache.put(f["hash"], enc)
            if enc is None:
                if hasattr(self.model, "embed_multimodal_item"):       # towers that need more than the pixels (grids …)
                    enc = self.model.embed_multimodal_item(f).to(emb.dtype)
                else:
                    px = torch.as_tensor(f["data"]).to(self.device)
                    enc = self.model.embed_multimodal(px[None])[0].to(emb.dtype)
                self.encoder_cache.put(f["hash"], enc)
                if self.ec_connector is not None:
                    self.ec_connector.save(f["hash"], enc)
            dst0 = int(cu[j]) + (lo - start)
            e0 = int(f.get("enc_offset", 0))                 # records that cover one slice of a shared encoder output
            rows_ = enc[e0 + lo - f["offset"]:e0 + hi - f["offset"]]
            Hd = emb.shape[-1]
            emb[dst0:dst0 + (hi - lo)] = rows_[:, :Hd]
            if rows_.shape[-1] > Hd:                         # towers that also emit per-layer features (DeepStack)
                if self._deepstack is None or self._deepstack.shape != (T, rows_.shape[-1] - Hd):
                    self._deepstack = torch.zeros(T, rows_.shape[-1] - Hd, dtype=emb.dtype, device=emb.device)
                    self._deepstack_live = True
                elif not self._deepstack_live:
                    self._deepstack.zero_()
                    self._deepstack_live = True
                self._deepstack[dst0:dst0 + (hi - lo)] = rows_[:, Hd:]
        return emb

    def _next_staging(self) -> _Staging:
        self._stage_i = (self._stage_i + 1) % len(self.staging)
        stg = self.staging[self._stage_i]
        t_w = time.perf_counter()
        stg.wait()  # the copies that used this set 2-3 steps ago have long completed
        self.wait_s += time.perf_counter() - t_w
        return stg

    def finalize_pending(self) -> None:
        if self._pending is not None:
            self._pending.get()

    @torch.inference_mode()
    def execute_model(self, so: SchedulerOutput, intermediate=None, non_block: bool = False):
        """Launch one step. Returns a ModelRunnerOutput, or (non_block) an AsyncStepOutput."""
        t0 = time.perf_counter()
        self._step += 1
        self._update_states(so)
        if so.total_num_scheduled_tokens == 0:
            if self.kv_connector is not None:            # requests may be waiting for remote KV: drive the transfers
                if so.kv_connector_metadata:
                    self.kv_connector.start_load_kv(so.kv_connector_metadata, self.kv_caches)
                return ModelRunnerOutput(kv_connector_output=self.kv_connector.get_finished())
            return ModelRunnerOutput()
        B_all = len(self.req_ids)
        # features that need the previous step's tokens on the host force a drain of the pipeline
        if self._pending is not None and (bool(self.complex_cpu[:B_all].any()) or so.scheduled_spec_decode_tokens
                                          or self.drafter is not None or so.grammar_bitmask is not None):
            self._pending.get()
        stg = self._next_staging()
        t_prep = time.perf_counter()
        gpu_prep = self.gpu_prep and not so.scheduled_spec_decode_tokens
        rows, ns, cu, T, max_q, max_s, n_fix = (self._prepare_gpu if gpu_prep else self._prepare)(so, stg)
        nb = len(rows)
        has_spec = bool(so.scheduled_spec_decode_tokens)
        decode_only = max_q == 1 and not has_spec
        pad_b = nb
        graph = None
        simple = not bool(self.complex_cpu[rows].any()) and so.grammar_bitmask is None
        no_filter = not bool(self.filter_cpu[rows].any())
        kv_meta = so.kv_connector_metadata if self.kv_connector is not None else None
        lora_batch = None
        if getattr(self, "lora_manager", None) is not None:
            lreqs = [self.states[self.req_ids[r]].lora_request for r in rows]
            lora_batch = self.lora_manager.make_batch(lreqs, [int(x) for x in ns])
        if (decode_only and self.use_cuda_graph and self.graphs and simple and no_filter and not kv_meta
                and lora_batch is None):
            pad_b = self._bucket(nb)
            graph = self.graphs.get(pad_b)
            if graph is None:
                pad_b = nb
        pad_t = pad_b if graph is not None else T
        mixed = None
        if (graph is None and not decode_only and self.mixed_graphs and gpu_prep and simple and no_filter and not kv_meta
                and lora_batch is None and not has_spec):
            for tb in sorted(self.mixed_graphs):
                if tb >= T:
                    if nb <= self.mixed_graphs[tb].b:
                        mixed = self.mixed_graphs[tb]
                    break
            if mixed is not None and any(self.states[self.req_ids[r]].mm_features for r in rows):
                mixed = None
        if mixed is not None:
            graph, pad_b, pad_t = mixed, mixed.b, mixed.t
        if gpu_prep:
            self._h2d_gpu(stg, nb, T, pad_b, max_s, pad_t if mixed is not None else 0)
        else:
            self._h2d(stg, rows, nb, T, pad_b, pad_t, max_s, n_fix)
        dt_prep = time.perf_counter() - t_prep
        self.prep_host_s += dt_prep
        if max_q == 1:                   # steady-state decode steps, admission / prefill excluded
            self.prep_decode_s += dt_prep
            self.prep_decode_n += 1
        cpu_qsl = stg.qsl[:nb + 1] if not self.is_cuda else None
        cpu_sl = stg.seq_lens[:nb] if not self.is_cuda else None
        # which rows produce a token this step (prefill chunk finished / decode)
        total_after = self.num_computed_cpu[rows] + ns
        produces = total_after >= self.num_logical_cpu[rows]  # spec rows: logical count excludes drafts
        lp_out, plp = None, {}
        req_ids = [self.req_ids[r] for r in rows]
        if graph is not None:
            # fast path: temperature / seed / step vectors straight from the per-row arrays
            stg.temperature[:nb] = torch.from_numpy(self.temp_cpu[rows])
            stg.seeds[:nb] = torch.from_numpy(self.seed_cpu[rows])
            stg.steps[:nb] = torch.from_numpy(self.num_logical_cpu[rows] - self.num_prompt_cpu[rows])
            if pad_b > nb:
                stg.temperature[nb:pad_b] = 0.0
            sampled = graph.replay_from_staging(stg, pad_b)[:nb]
            self.launch_count += graph.num_launches
            self.h2d_bytes += pad_b * 16
        else:
            l0 = self._launches()
            states = [self.states[r] for r in req_ids]
            md = self._make_metadata(nb, T, max_q, max_s, decode_only, cpu_qsl, cpu_sl)
            if so.num_common_prefix_blocks and nb >= self.cascade_min_reqs and self.pcp == 1 and self.dcp == 1:
                self._add_cascade(md, so.num_common_prefix_blocks, rows, ns, nb, T, max_q, max_s)
            if kv_meta:
                self.kv_connector.start_load_kv(kv_meta, self.kv_caches)     # external prefix → our blocks
            extra = None
            if getattr(self.drafter, "aux_layers", None):
                extra = {"aux_layers": self.drafter.aux_layers}
            if self.is_enc_dec:
                extra = {"cross": self._cross_context(states, cu)}
                mm_emb = None
            else:
                mm_emb = (self._mm_inputs_embeds(states, rows, ns, cu, T)
                          if self.has_mm or any(s.mm_features for s in states) else None)
                if self.uses_mrope:
                    mpos = self._mrope_positions(states, rows, ns, cu, T)
                    if mpos is not None:
                        extra = dict(extra or {}, mrope_positions=mpos)
                if mm_emb is not None and self._deepstack_live:
                    extra = dict(extra or {}, deepstack=self._deepstack)
            pcp_last = None
            if self.pcp > 1:
                if (lora_batch is not None or mm_emb is not None or extra or has_spec or self.drafter is not None
                        or any(s.pooling_params is not None or (s.sampling_params is not None
                               and s.sampling_params.prompt_logprobs is not None) for s in states)):
                    raise NotImplementedError("prefill context parallelism serves plain text generation only "
                                              "(no LoRA / multimodal / speculative decoding / pooling / prompt logprobs)")
                hidden = pcp_last = self._forward_pcp(T, nb, cu, ns, (self.num_computed_cpu[rows] + ns).astype(np.int64),
                                                      intermediate)
            elif lora_batch is not None:
                from vllm_b200 import lora as _lora
                _lora.set_batch(lora_batch)
                try:
                    hidden = self._forward(T, md, intermediate, mm_emb, extra)
                finally:
                    _lora.set_batch(None)
            else:
                hidden = self._forward(T, md, intermediate, mm_emb, extra)
            if kv_meta:
                self.kv_connector.save_kv(kv_meta, self.kv_caches)           # freshly computed full blocks
            if isinstance(hidden, dict):  # non-last PP stage
                return ModelRunnerOutput(kv_connector_output=hidden)
            if any(s.pooling_params is not None for s in states):
                return self._pool(hidden, states, rows, cu, produces, t0)
            if has_spec:
                return self._verify_spec(so, hidden, states, rows, ns, cu, t0)
            last_idx = torch.from_numpy(cu[1:].astype(np.int64) - 1).to(self.device, non_blocking=True)
            if getattr(self.drafter, "needs_step_hidden", False):
                # EAGLE / MTP heads consume the target hidden state of every position computed this step
                src = self._last_aux if getattr(self.drafter, "aux_layers", None) else hidden
                self.step_hidden = (src, rows.copy(), cu.copy(), self.num_computed_cpu[rows].copy())
            last_hidden = hidden[last_idx] if pcp_last is None else pcp_last
            if self.spec_hidden is not None:        # hidden-state proposers (Medusa) read the newest position's state
                self.spec_hidden.index_copy_(0, torch.from_numpy(rows.astype(np.int64)).to(self.device), last_hidden)
            logits = self.model.compute_logits(last_hidden)
            smd = self._sampling_metadata(states, rows, nb, simple)
            bitmask = self._gather_bitmask(so, states) if so.grammar_bitmask is not None else None
            so_out = self.sampler(logits, smd, bitmask)
            sampled = so_out.sampled_token_ids
            if so_out.logprobs is not None:
                keep = torch.from_numpy(np.nonzero(produces)[0]).to(self.device)
                ids, vals, ranks = so_out.logprobs
                lp_out = LogprobsTensors(ids[keep].cpu(), vals[keep].cpu(), ranks[keep].cpu())
            if not simple and any(s.sampling_params is not None and s.sampling_params.prompt_logprobs is not None
                                  for s in states):
                plp = self._prompt_logprobs(hidden, states, rows, ns, cu)
            if self.is_cuda:
                self.launch_count += self._launches() - l0
        # publish: device-side last_sampled (for the next step's input patch) + async D2H
        ev = None
        if self.is_cuda:
            rows_dev = stg.rows[:nb].to(self.device, non_blocking=True)
            self.last_sampled.index_copy_(0, rows_dev, sampled)
            if gpu_prep:        # the new ids become next step's inputs on the device (no D2H → H2D round trip)
                self._cu.append_sampled(self.dev_tokens, self.step_meta[:nb], self.step_meta[3 * nb:4 * nb], sampled)
            stg.sampled[:nb].copy_(sampled, non_blocking=True)
            stg.event.record()
            ev = stg.event
        else:
            stg.sampled[:nb] = sampled
        self.num_logical_cpu[rows] += produces.astype(np.int32)
        self.num_computed_cpu[rows] = total_after
        if self.eplb_interval and self._step % self.eplb_interval == 0 and self.eplb_hook is not None:
            self.eplb_hook()              # every EP rank reaches the same step count: the shuffle is a collective
        out = AsyncStepOutput(self, req_ids, produces, stg, nb, ev, t0, lp_out, plp)
        if non_block and self.async_mode:
            self._pending = out
            return out
        return out.get()

    # ------------------------------------------------------------- helpers
    def _launches(self) -> int:
        if not self.is_cuda:
            return 0
        from vllm_b200.ops import cuda_ops
        return cuda_ops().LAUNCHES

    def _bucket(self, n: int) -> int:
        for s in self.capture_sizes:
            if s >= n:
                return s
        return -1

    def _sampling_metadata(self, states: list[ReqState], rows, pad_b: int, simple: bool = False) -> SamplingMetadata:
        B = len(states)
        dev = self.device
        rows = np.asarray(rows)
        temps = self.temp_cpu[rows].copy()
        all_greedy = bool((temps < 1e-5).all())
        all_random = bool((temps >= 1e-5).all())
        t = torch.from_numpy(temps).to(dev, non_blocking=True)
        if simple and not bool(self.filter_cpu[rows].any()):
            md = SamplingMetadata(temperature=t, all_greedy=all_greedy, all_random=all_random)
            if not all_greedy:
                md.seeds = torch.from_numpy(self.seed_cpu[rows]).to(dev)
                md.steps = torch.from_numpy((self.num_logical_cpu[rows] - self.num_prompt_cpu[rows]).astype(np.int32)).to(dev)
                if not self.is_cuda:  # the reference sampler draws from per-request torch generators
                    md.generators = {i: s.generator for i, s in enumerate(states) if s.generator is not None}
            return md
        top_k = np.zeros(B, dtype=np.int32)
        top_p = np.ones(B, dtype=np.float32)
        min_p = np.zeros(B, dtype=np.float32)
        pres = np.zeros(B, dtype=np.float32)
        freq = np.zeros(B, dtype=np.float32)
        rep = np.ones(B, dtype=np.float32)
        gens: dict[int, torch.Generator] = {}
        logit_bias: list = [None] * B
        bad: dict[int, list[list[int]]] = {}
        min_toks: dict[int, tuple[int, set[int]]] = {}
        max_lp: Optional[int] = None
        allowed_rows = []
        any_pen = False
        for i, s in enumerate(states):
            sp = s.sampling_params
            if sp is None:
                continue
            top_k[i], top_p[i], min_p[i] = sp.top_k, sp.top_p, sp.min_p
            pres[i], freq[i], rep[i] = sp.presence_penalty, sp.frequency_penalty, sp.repetition_penalty
            any_pen |= sp.needs_penalties
            if s.generator is not None:
                gens[i] = s.generator
            if sp.logit_bias:
                logit_bias[i] = sp.logit_bias
            if sp.bad_words_token_ids:
                bad[i] = sp.bad_words_token_ids
            if sp.min_tokens > 0 and len(s.output_token_ids) < sp.min_tokens:
                min_toks[i] = (sp.min_tokens, set(sp.all_stop_token_ids))
            if sp.logprobs is not None:
                k = self.vocab_size if sp.logprobs == -1 else sp.logprobs
                max_lp = k if max_lp is None else max(max_lp, k)
            if sp.allowed_token_ids:
                allowed_rows.append(i)
        md = SamplingMetadata(temperature=t, all_greedy=all_greedy, all_random=all_random, generators=gens,
                              max_num_logprobs=max_lp, logit_bias=logit_bias if any(logit_bias) else [],
                              bad_words_token_ids=bad, min_tokens=min_toks,
                              output_token_ids=[s.output_token_ids for s in states])
        if self.sampler.custom:
            md.extra_args = [s.sampling_params.extra_args if s.sampling_params is not None else None for s in states]
            md.prompt_token_ids = [self.token_ids_cpu[s.row, :s.num_prompt_tokens].tolist() for s in states]
        if not all_greedy:
            if (top_k > 0).any():
                md.top_k = torch.from_numpy(top_k).to(dev)
            if (top_p < 1.0).any():
                md.top_p = torch.from_numpy(top_p).to(dev)
            if (min_p > 0).any():
                md.min_p = torch.from_numpy(min_p).to(dev)
            md.seeds = torch.from_numpy(self.seed_cpu[rows]).to(dev)
            md.steps = torch.from_numpy((self.num_logical_cpu[rows] - self.num_prompt_cpu[rows]).astype(np.int32)).to(dev)
        if any_pen:
            md.no_penalties = False
            md.presence, md.frequency = torch.from_numpy(pres).to(dev), torch.from_numpy(freq).to(dev)
            md.repetition = torch.from_numpy(rep).to(dev)
            V = self.vocab_size
            pm = torch.zeros(B, V, dtype=torch.bool)
            oc = torch.zeros(B, V, dtype=torch.int32)
            for i, s in enumerate(states):
                r = s.row
                p = torch.from_numpy(self.token_ids_cpu[r, :s.num_prompt_tokens].astype(np.int64))
                pm[i, p.clamp(max=V - 1)] = True
                if s.output_token_ids:
                    o = torch.tensor(s.output_token_ids, dtype=torch.int64).clamp(max=V - 1)
                    oc[i].index_add_(0, o, torch.ones_like(o, dtype=torch.int32))
            md.prompt_mask, md.output_counts = pm.to(dev), oc.to(dev)
        if allowed_rows:
            m = torch.zeros(B, self.vocab_size, dtype=torch.bool)
            for i in allowed_rows:
                m[i] = True
                m[i, torch.tensor(states[i].sampling_params.allowed_token_ids)] = False
            md.allowed_token_ids_mask = m.to(dev)
        return md

    def _gather_bitmask(self, so: SchedulerOutput, states: list[ReqState], expand: bool = False
                        ) -> Optional[torch.Tensor]:
        """Grammar bitmask rows in batch order.  The scheduler lays the mask out per structured request with one row
        per verified position (k drafts → k + 1 rows); ``expand`` returns all of them (spec verification: one row
        per target-logit row), otherwise the first row of each request (plain sampling)."""
        bm = so.grammar_bitmask
        spec = so.scheduled_spec_decode_tokens
        start, off = {}, 0
        for rid in so.structured_output_request_ids:
            start[rid] = off
            off += len(spec.get(rid) or ()) + 1
        W = bm.shape[1]
        n_rows = sum(len(spec.get(s.req_id) or ()) + 1 for s in states) if expand else len(states)
        full = np.full((n_rows, W), -1, dtype=np.int32)
        r = 0
        for s in states:
            k = (len(spec.get(s.req_id) or ()) + 1) if expand else 1
            i = start.get(s.req_id)
            if i is not None:
                full[r:r + k] = bm[i:i + k]
            r += k
        return torch.from_numpy(full).to(self.device)

    def _prompt_logprobs(self, hidden, states, rows, ns, cu) -> dict[str, Optional[LogprobsTensors]]:
        out: dict[str, Optional[LogprobsTensors]] = {}
        for j, s in enumerate(states):
            sp = s.sampling_params
            if sp is None or sp.prompt_logprobs is None:
                continue
            r = rows[j]
            start = int(self.num_computed_cpu[r])
            if start >= s.num_prompt_tokens - 1:
                continue
            n = min(int(ns[j]), s.num_prompt_tokens - 1 - start)
            if n <= 0:
                continue
            h = hidden[cu[j]:cu[j] + n]
            logits = self.model.compute_logits(h).float()
            tgt = torch.from_numpy(self.token_ids_cpu[r, start + 1:start + 1 + n].astype(np.int64)).to(self.device)
            k = self.vocab_size if sp.prompt_logprobs == -1 else sp.prompt_logprobs
            from vllm_b200.ops import ref
            ids, vals, ranks = ref.compute_logprobs(logits, tgt, k)
            out[s.req_id] = LogprobsTensors(ids.cpu(), vals.cpu(), ranks.cpu())
        return out

    def _pool(self, hidden, states, rows, cu, produces, t0) -> ModelRunnerOutput:
        from vllm_b200.layers.pooler import pool_hidden
        req_ids, pooled, out_ids = [], [], []
        for j, s in enumerate(states):
            r = rows[j]
            req_ids.append(s.req_id)
            out_ids.append([])
            self.num_computed_cpu[r] = self.num_computed_cpu[r] + (cu[j + 1] - cu[j])
            if produces[j] and s.pooling_params is not None:
                pooled.append(pool_hidden(hidden[cu[j]:cu[j + 1]], s.pooling_params, self.model).cpu())
            else:
                pooled.append(None)
        return ModelRunnerOutput(req_ids=req_ids, req_id_to_index={r: i for i, r in enumerate(req_ids)},
                                 sampled_token_ids=out_ids, pooler_output=pooled,
                                 forward_ms=(time.perf_counter() - t0) * 1e3)

    def _verify_spec(self, so, hidden, states, rows, ns, cu, t0) -> ModelRunnerOutput:
        """Target-model verification of scheduled draft tokens."""
        idx_list, draft_lists, cu_l = [], [], [0]
        for j, s in enumerate(states):
            drafts = so.scheduled_spec_decode_tokens.get(s.req_id, [])
            k = len(drafts)
            end = int(cu[j + 1])
            idx_list.extend(range(end - k - 1, end))
            draft_lists.append(list(drafts))
            cu_l.append(cu_l[-1] + k + 1)
        if getattr(self.drafter, "needs_step_hidden", False):
            src = self._last_aux if getattr(self.drafter, "aux_layers", None) else hidden
            self.step_hidden = (src, rows.copy(), cu.copy(), self.num_computed_cpu[rows].copy())
        idx = torch.tensor(idx_list, dtype=torch.int64, device=self.device)
        logits = self.model.compute_logits(hidden[idx]).float()
        # one metadata row per verified position: the request's sampling state as it will be AFTER drafts[:p], so the
        # grammar mask, processors, penalties and top-k/p filters of the plain sampler apply to every position
        import dataclasses
        exp_states, exp_rows = [], []
        for j, s in enumerate(states):
            for pos in range(len(draft_lists[j]) + 1):
                exp_states.append(s if pos == 0 else
                                  dataclasses.replace(s, output_token_ids=s.output_token_ids + draft_lists[j][:pos]))
                exp_rows.append(rows[j])
        smd = self._sampling_metadata(exp_states, exp_rows, len(exp_states))
        bitmask = self._gather_bitmask(so, states, expand=True) if so.grammar_bitmask is not None else None
        total_after = self.num_computed_cpu[rows] + ns
        accepted = self.rejection_sampler(logits, draft_lists, smd, cu_l, bitmask)
        lp_tensors = None
        if smd.max_num_logprobs is not None:
            sel, toks = [], []
            for j, acc in enumerate(accepted):
                if total_after[j] - len(draft_lists[j]) < self.num_tokens_cpu[rows[j]]:
                    continue                      # still prefilling: nothing is emitted for this request
                sel.extend(range(cu_l[j], cu_l[j] + len(acc)))
                toks.extend(acc)
            if sel:
                si = torch.tensor(sel, dtype=torch.int64, device=self.device)
                from vllm_b200.ops import ref as ref_ops
                lp = ref_ops.compute_logprobs(logits[si], torch.tensor(toks, dtype=torch.int64, device=self.device),
                                              smd.max_num_logprobs)
                lp_tensors = LogprobsTensors(lp[0].cpu(), lp[1].cpu(), lp[2].cpu())
        req_ids, out_ids = [], []
        for j, s in enumerate(states):
            r = rows[j]
            req_ids.append(s.req_id)
            k = len(draft_lists[j])
            known = self.num_tokens_cpu[r]
            if total_after[j] - k < known:  # still prefilling
                out_ids.append([])
                self.num_computed_cpu[r] = total_after[j]
                continue
            acc = accepted[j]
            n = known
            self.token_ids_cpu[r, n:n + len(acc)] = acc
            self.num_tokens_cpu[r] = n + len(acc)
            self.num_logical_cpu[r] = n + len(acc)
            s.output_token_ids.extend(acc)
            out_ids.append(acc)
            # rejected drafts' KV slots are simply overwritten next step
            self.num_computed_cpu[r] = total_after[j] - (k + 1 - len(acc))
            if self.spec_hidden is not None and acc:
                # state of the position that produced the last emitted token
                self.spec_hidden[r] = hidden[int(cu[j + 1]) - k - 1 + len(acc) - 1]
        spec_ids = self.drafter.propose(self, rows, out_ids) if self.drafter is not None else None
        return ModelRunnerOutput(req_ids=req_ids, req_id_to_index={r: i for i, r in enumerate(req_ids)},
                                 sampled_token_ids=out_ids, spec_token_ids=spec_ids, logprobs=lp_tensors,
                                 forward_ms=(time.perf_counter() - t0) * 1e3)

    # --------------------------------------------------------- profiling
    @torch.inference_mode()
    def profile_run(self) -> None:
        """Dummy forward at ``max_num_batched_tokens`` with no KV cache (peak activation memory)."""
        T = min(self.max_num_tokens, self.max_model_len * self.max_num_reqs)
        self.input_ids[:T].zero_()
        self.positions[:T].zero_()
        hidden = self._forward(T, {})
        if not isinstance(hidden, dict) and not getattr(self.model, "is_pooling_model", False):
            n = min(self.max_num_reqs, T)
            logits = self.model.compute_logits(hidden[:n])
            logits.float().argmax(-1)
        if self.is_cuda:
            torch.cuda.synchronize()

    # -------------------------------------------------------- CUDA graphs
    def _attention_kernels_cover_model(self) -> bool:
        """False when some attention layer would fall back to the plain-PyTorch op (which syncs with the host)."""
        from vllm_b200.ops import cuda_attention_supported
        for name, cache in self.kv_caches.items():
            layer = attention_layers().get(name)
            if layer is None or cache.dim() != 5:
                continue
            if not cuda_attention_supported(cache.shape[4], cache.shape[3], self.dtype, self.kv_vt[self.layer_to_group[name]]):
                return False
        return True

    def capture_model(self) -> None:
        if not self.use_cuda_graph or getattr(self.model, "is_pooling_model", False):
            return
        if not self._attention_kernels_cover_model():
            logger.warning("attention shapes outside the CUDA kernels' coverage: running eagerly on the reference ops")
            self.use_cuda_graph = False
            return
        from vllm_b200.worker.cudagraph import DecodeGraph
        sizes = [b for b in self.vllm_config.compilation_config.capture_sizes(self.max_num_reqs) if b <= self.max_num_reqs]
        self.capture_sizes = sorted(sizes)
        t0 = time.perf_counter()
        try:
            for b in reversed(self.capture_sizes):
                self.graphs[b] = DecodeGraph(self, b)
        except Exception as e:      # a layer fell back to an op that cannot be captured (host sync, unpinned copy)
            logger.warning("CUDA-graph capture failed (%s: %s): decode steps run eagerly", type(e).__name__, str(e)[:200])
            self.graphs.clear()
            self.use_cuda_graph = False
            torch.cuda.synchronize()
            return
        torch.cuda.synchronize()
        logger.info("captured %d decode CUDA graphs in %.1fs", len(self.graphs), time.perf_counter() - t0)
        self._capture_mixed()

    def _capture_mixed(self) -> None:
        """Mixed prefill+decode buckets (cudagraph_mode FULL): needs the device-side input expansion and the tensor-core
        attention layout on every KV group (the flat varlen schedule lives in that kernel)."""
        sizes = self.vllm_config.compilation_config.mixed_sizes(self.max_num_tokens)
        if not sizes or not self.gpu_prep or not all(self.kv_vt) or self.has_mm or self.drafter is not None:
            return
        from vllm_b200.worker.cudagraph import MixedGraph
        t0 = time.perf_counter()
        try:
            for t in reversed(sizes):
                self.mixed_graphs[t] = MixedGraph(self, t, min(t, self.max_num_reqs, 256))
        except Exception as e:
            logger.warning("mixed-batch CUDA-graph capture failed (%s: %s): mixed steps run eagerly", type(e).__name__,
                           str(e)[:200])
            self.mixed_graphs.clear()
        torch.cuda.synchronize()
        if self.mixed_graphs:
            logger.info("captured %d mixed-batch CUDA graphs %s in %.1fs", len(self.mixed_graphs),
                        sorted(self.mixed_graphs), time.perf_counter() - t0)
