"""Offline inference API.

Behavioural reference: vllm/entrypoints/llm.py (``LLM`` :67; ``generate`` :418,
``chat`` :612, ``encode``/``embed``/``classify``/``score``, ``beam_search``,
``collective_rpc`` :567, ``sleep``/``wake_up`` :803,828), offline_utils.py
``_run_engine`` :573-626.
"""
from __future__ import annotations

import itertools
from typing import Any, Optional, Sequence, Union

from vllm_b200.engine.arg_utils import EngineArgs
from vllm_b200.engine.llm_engine import LLMEngine
from vllm_b200.inputs import PromptType
from vllm_b200.outputs import (BeamSearchOutput, BeamSearchSequence, EmbeddingRequestOutput, PoolingRequestOutput,
                               RequestOutput)
from vllm_b200.sampling_params import BeamSearchParams, PoolingParams, RequestOutputKind, SamplingParams


class LLM:
    def __init__(self, model: str, **kwargs: Any) -> None:
        mp = kwargs.pop("multiprocess_mode", None)
        kwargs.setdefault("disable_log_stats", True)
        self.engine_args = EngineArgs(model=model, **kwargs)
        self.llm_engine = LLMEngine.from_engine_args(self.engine_args, multiprocess_mode=mp)
        self.request_counter = itertools.count()
        self.default_sampling_params: Optional[dict] = None

    def get_tokenizer(self):
        return self.llm_engine.get_tokenizer()

    def get_default_sampling_params(self) -> SamplingParams:
        return SamplingParams()

    # ---------------------------------------------------------------- generate
    def generate(self, prompts: Union[PromptType, Sequence[PromptType]],
                 sampling_params: Union[SamplingParams, Sequence[SamplingParams], None] = None,
                 use_tqdm: bool = False, lora_request=None, priority: Optional[list[int]] = None
                 ) -> list[RequestOutput]:
        prompts = self._as_list(prompts)
        if sampling_params is None:
            sampling_params = self.get_default_sampling_params()
        self._add(prompts, sampling_params, lora_request, priority, final_only=True)
        return self._run_engine(use_tqdm)

    def enqueue(self, prompts, sampling_params=None, lora_request=None, priority=None) -> list[str]:
        prompts = self._as_list(prompts)
        return self._add(prompts, sampling_params or self.get_default_sampling_params(), lora_request, priority,
                         final_only=True)

    def wait_for_completion(self, use_tqdm: bool = False) -> list[RequestOutput]:
        return self._run_engine(use_tqdm)

    @staticmethod
    def _as_list(prompts) -> list:
        if isinstance(prompts, (str, dict)):
            return [prompts]
        if isinstance(prompts, list) and prompts and isinstance(prompts[0], int):
            return [prompts]
        return list(prompts)

    def _add(self, prompts, params, lora_request, priority, final_only: bool) -> list[str]:
        if isinstance(params, (list, tuple)) and len(params) != len(prompts):
            raise ValueError("the lengths of prompts and sampling_params must match")
        ids = []
        for i, p in enumerate(prompts):
            sp = params[i] if isinstance(params, (list, tuple)) else params
            if isinstance(sp, SamplingParams) and final_only:
                sp = sp.clone()
                sp.output_kind = RequestOutputKind.FINAL_ONLY
            rid = str(next(self.request_counter))
            lr = lora_request[i] if isinstance(lora_request, (list, tuple)) else lora_request
            self.llm_engine.add_request(rid, p, sp, lora_request=lr, priority=priority[i] if priority else 0)
            ids.append(rid)
        return ids

    def _run_engine(self, use_tqdm: bool = False) -> list:
        outs = []
        pbar = None
        if use_tqdm:
            from tqdm import tqdm
            pbar = tqdm(total=self.llm_engine.get_num_unfinished_requests(), desc="Processed prompts")
        while self.llm_engine.has_unfinished_requests():
            try:
                step_outs = self.llm_engine.step()
            except BaseException:
                # leave the engine reusable: drop every request of this call (front-end state and scheduler)
                try:
                    self.llm_engine.abort_request(list(self.llm_engine.output_processor.request_states))
                except Exception:
                    pass
                raise
            for o in step_outs:
                if o.finished:
                    outs.append(o)
                    if pbar:
                        pbar.update(1)
        if pbar:
            pbar.close()
        return sorted(outs, key=lambda o: int(o.request_id) if str(o.request_id).isdigit() else 0)

    # -------------------------------------------------------------------- chat
    def chat(self, messages: Union[list[dict], list[list[dict]]], sampling_params=None, use_tqdm: bool = False,
             chat_template: Optional[str] = None, add_generation_prompt: bool = True,
             continue_final_message: bool = False, tools: Optional[list[dict]] = None, lora_request=None,
             chat_template_kwargs: Optional[dict] = None) -> list[RequestOutput]:
        from vllm_b200.entrypoints.chat_utils import chat_multi_modal_data, render_chat
        convs = messages if messages and isinstance(messages[0], list) else [messages]
        tok = self.get_tokenizer()
        prompts = []
        for c in convs:
            text = render_chat(tok, c, chat_template, add_generation_prompt, continue_final_message, tools,
                               **(chat_template_kwargs or {}))
            mm = chat_multi_modal_data(c)
            prompts.append({"prompt": text, "multi_modal_data": mm} if mm else text)
        return self.generate(prompts, sampling_params, use_tqdm, lora_request)

    # ----------------------------------------------------------------- pooling
    def encode(self, prompts, pooling_params: Optional[PoolingParams] = None, use_tqdm: bool = False,
               pooling_task: str = "embed") -> list[PoolingRequestOutput]:
        prompts = self._as_list(prompts)
        pp = pooling_params or PoolingParams(task=pooling_task)
        self._add(prompts, pp, None, None, final_only=False)
        return self._run_engine(use_tqdm)

    def embed(self, prompts, pooling_params=None, use_tqdm: bool = False) -> list[EmbeddingRequestOutput]:
        outs = self.encode(prompts, pooling_params or PoolingParams(task="embed"), use_tqdm)
        return [EmbeddingRequestOutput(o.request_id, o.outputs, o.prompt_token_ids, o.finished) for o in outs]

    def classify(self, prompts, use_tqdm: bool = False):
        return self.encode(prompts, PoolingParams(task="classify", normalize=False), use_tqdm)

    def score(self, text_1, text_2, use_tqdm: bool = False) -> list[float]:
        """Cosine similarity of embeddings (bi-encoder scoring)."""
        import torch
        a = self._as_list(text_1)
        b = self._as_list(text_2)
        if len(a) == 1:
            a = a * len(b)
        ea = [torch.as_tensor(o.outputs.data).float() for o in self.encode(a)]
        eb = [torch.as_tensor(o.outputs.data).float() for o in self.encode(b)]
        return [float(torch.nn.functional.cosine_similarity(x, y, dim=0)) for x, y in zip(ea, eb)]

    # -------------------------------------------------------------- beam search
    def beam_search(self, prompts, params: BeamSearchParams) -> list[BeamSearchOutput]:
        tok = self.get_tokenizer()
        results = []
        for p in self._as_list(prompts):
            if isinstance(p, dict) and "prompt_token_ids" in p:
                ids = list(p["prompt_token_ids"])
            else:
                ids = tok.encode(p if isinstance(p, str) else p["prompt"])
            beams = [BeamSearchSequence(tokens=ids, cum_logprob=0.0)]
            done: list[BeamSearchSequence] = []
            eos = self.llm_engine.model_config.eos_token_id
            for _ in range(params.max_tokens):
                sp = SamplingParams(max_tokens=1, logprobs=2 * params.beam_width, temperature=params.temperature,
                                    detokenize=False)
                outs = self.generate([{"prompt_token_ids": b.tokens} for b in beams], sp)
                cand = []
                for b, o in zip(beams, outs):
                    lp = o.outputs[0].logprobs[0]
                    for tid, l in lp.items():
                        nb = BeamSearchSequence(tokens=b.tokens + [tid], cum_logprob=b.cum_logprob + l.logprob,
                                                logprobs=b.logprobs + [lp])
                        if tid == eos and not params.ignore_eos:
                            nb.finish_reason = "stop"
                            done.append(nb)
                        else:
                            cand.append(nb)
                n_prompt = len(ids)
                score = lambda s: s.cum_logprob / ((len(s.tokens) - n_prompt) ** params.length_penalty)
                beams = sorted(cand, key=score, reverse=True)[:params.beam_width]
                if not beams:
                    break
            done.extend(beams)
            n_prompt = len(ids)
            best = sorted(done, key=lambda s: s.cum_logprob / max(len(s.tokens) - n_prompt, 1) ** params.length_penalty,
                          reverse=True)[:params.beam_width]
            for s in best:
                s.text = tok.decode(s.tokens[n_prompt:]) if tok is not None else None
            results.append(BeamSearchOutput(best))
        return results

    # --------------------------------------------------------------------- misc
    def collective_rpc(self, method, args=(), kwargs=None):
        return self.llm_engine.collective_rpc(method, args, kwargs)

    def start_profile(self):
        self.llm_engine.start_profile()

    def stop_profile(self):
        self.llm_engine.stop_profile()

    def reset_prefix_cache(self):
        return self.llm_engine.reset_prefix_cache()

    def sleep(self, level: int = 1):
        self.reset_prefix_cache()
        self.llm_engine.sleep(level)

    def wake_up(self, tags=None):
        self.llm_engine.wake_up(tags)

    def get_metrics(self):
        return self.llm_engine.get_metrics()

    def shutdown(self):
        self.llm_engine.shutdown()
