"""Architecture registry (reference: vllm/model_executor/models/registry.py).

Maps HF ``architectures[0]`` to ``module:Class`` lazily so importing the package
does not import every model file.
"""
from __future__ import annotations

import importlib
from typing import Type

import torch.nn as nn

_MODELS: dict[str, tuple[str, str]] = {
    "LlamaForCausalLM": ("llama", "LlamaForCausalLM"),
    "LLaMAForCausalLM": ("llama", "LlamaForCausalLM"),
    "MistralForCausalLM": ("llama", "MistralForCausalLM"),
    "Qwen2ForCausalLM": ("llama", "Qwen2ForCausalLM"),
    "Qwen3ForCausalLM": ("llama", "Qwen3ForCausalLM"),
    "GemmaForCausalLM": ("llama", "GemmaForCausalLM"),
    "Phi3ForCausalLM": ("llama", "Phi3ForCausalLM"),
    "GraniteMoeForCausalLM": ("granite_moe", "GraniteMoeForCausalLM"),
    "GraniteMoeSharedForCausalLM": ("granite_moe", "GraniteMoeForCausalLM"),
    "GraniteMoeHybridForCausalLM": ("granite_moe", "GraniteMoeForCausalLM"),
    "GraniteForCausalLM": ("llama_variants", "GraniteForCausalLM"),
    "Olmo2ForCausalLM": ("llama_variants", "Olmo2ForCausalLM"),
    "CohereForCausalLM": ("llama_variants", "CohereForCausalLM"),
    "PhiForCausalLM": ("llama_variants", "PhiForCausalLM"),
    "Glm4ForCausalLM": ("llama_variants", "Glm4ForCausalLM"),
    "NemotronForCausalLM": ("llama_variants", "NemotronForCausalLM"),
    "Gemma3ForCausalLM": ("llama_variants", "Gemma3ForCausalLM"),
    "SmolLM3ForCausalLM": ("llama_variants", "SmolLM3ForCausalLM"),
    "Cohere2ForCausalLM": ("llama_variants", "Cohere2ForCausalLM"),
    "Olmo3ForCausalLM": ("llama_variants", "Olmo3ForCausalLM"),
    "Exaone4ForCausalLM": ("llama_variants", "Exaone4ForCausalLM"),
    "SeedOssForCausalLM": ("llama_variants", "SeedOssForCausalLM"),
    "ArceeForCausalLM": ("llama_variants", "ArceeForCausalLM"),
    "HeliumForCausalLM": ("llama_variants", "HeliumForCausalLM"),
    "Ernie4_5ForCausalLM": ("llama_variants", "Ernie4_5ForCausalLM"),
    "OlmoForCausalLM": ("llama_variants", "OlmoForCausalLM"),
    "StableLmForCausalLM": ("llama_variants", "StableLmForCausalLM"),
    "StableLMEpochForCausalLM": ("llama_variants", "StableLmForCausalLM"),
    "MinistralForCausalLM": ("llama_variants", "MinistralForCausalLM"),
    "GlmForCausalLM": ("llama_variants", "GlmForCausalLM"),
    "HunYuanDenseV1ForCausalLM": ("llama_variants", "HunYuanDenseV1ForCausalLM"),
    "VaultGemmaForCausalLM": ("llama_variants", "VaultGemmaForCausalLM"),
    "HunYuanMoEV1ForCausalLM": ("llama_variants", "HunYuanMoEV1ForCausalLM"),
    "PhimoeForCausalLM": ("llama_variants", "PhimoeForCausalLM"),
    "PhiMoEForCausalLM": ("llama_variants", "PhimoeForCausalLM"),
    "ApertusForCausalLM": ("llama_variants", "ApertusForCausalLM"),
    "CwmForCausalLM": ("llama_variants", "CwmForCausalLM"),
    "Jais2ForCausalLM": ("llama_variants", "Jais2ForCausalLM"),
    "Ministral3ForCausalLM": ("llama_variants", "Ministral3ForCausalLM"),
    "MistralLarge3ForCausalLM": ("llama_variants", "Ministral3ForCausalLM"),
    "FlexOlmoForCausalLM": ("llama_variants", "FlexOlmoForCausalLM"),
    "Glm4MoeLiteForCausalLM": ("deepseek", "Glm4MoeLiteForCausalLM"),
    "GlmMoeDsaForCausalLM": ("deepseek", "GlmMoeDsaForCausalLM"),
    "DeepseekV32ForCausalLM": ("deepseek", "DeepseekV32ForCausalLM"),
    "DbrxForCausalLM": ("llama_variants", "DbrxForCausalLM"),
    "AfmoeForCausalLM": ("llama_variants", "AfmoeForCausalLM"),
    "MiniMaxM2ForCausalLM": ("llama_variants", "MiniMaxM2ForCausalLM"),
    "ExaoneMoeForCausalLM": ("llama_variants", "ExaoneMoeForCausalLM"),
    "SolarOpenForCausalLM": ("deepseek", "SolarOpenForCausalLM"),
    "InternLM2ForCausalLM": ("llama_aliases", "InternLM2ForCausalLM"),
    "InternLM3ForCausalLM": ("llama_aliases", "InternLM3ForCausalLM"),
    "IQuestCoderForCausalLM": ("llama", "LlamaForCausalLM"),
    "TeleChat3ForCausalLM": ("llama", "LlamaForCausalLM"),
    "ExaoneForCausalLM": ("llama_aliases", "ExaoneForCausalLM"),
    "MiMoForCausalLM": ("llama_aliases", "MiMoForCausalLM"),
    "MiniCPM3ForCausalLM": ("deepseek", "MiniCPM3ForCausalLM"),
    "MiniCPMForCausalLM": ("llama_aliases", "MiniCPMForCausalLM"),
    "TeleFLMForCausalLM": ("llama_aliases", "TeleFLMForCausalLM"),
    "OrionForCausalLM": ("llama_aliases", "OrionForCausalLM"),
    "GraniteSWAForCausalLM": ("llama_variants", "GraniteForCausalLM"),
    "GraniteMoeSWAForCausalLM": ("granite_moe", "GraniteMoeForCausalLM"),
    "DeepseekForCausalLM": ("deepseek", "DeepseekForCausalLM"),
    "TransformersForCausalLM": ("transformers_backend", "TransformersForCausalLM"),
    "TransformersMoEForCausalLM": ("transformers_backend", "TransformersForCausalLM"),
    "RobertaForMaskedLM": ("bert", "RobertaModel"),
    "NomicBertModel": ("nomic_bert", "NomicBertModel"),
    "BgeM3EmbeddingModel": ("bert", "BgeM3EmbeddingModel"),
    "HF_ColBERT": ("bert", "HF_ColBERT"),
    "ColBERTModernBertModel": ("modernbert", "ColBERTModernBertModel"),
    "ColBERTJinaRobertaModel": ("bert", "ColBERTJinaRobertaModel"),
    "BertSpladeSparseEmbeddingModel": ("bert", "BertSpladeSparseEmbeddingModel"),
    "RobertaForSequenceClassification": ("bert", "RobertaForSequenceClassification"),
    "XLMRobertaForSequenceClassification": ("bert", "RobertaForSequenceClassification"),
    "BertForTokenClassification": ("bert", "BertForTokenClassification"),
    "RobertaForTokenClassification": ("bert", "RobertaForTokenClassification"),
    "XLMRobertaForTokenClassification": ("bert", "RobertaForTokenClassification"),
    "ModernBertForSequenceClassification": ("modernbert", "ModernBertForSequenceClassification"),
    "ModernBertForTokenClassification": ("modernbert", "ModernBertForTokenClassification"),
    "LlamaBidirectionalModel": ("llama_aliases", "LlamaBidirectionalModel"),
    "LlamaBidirectionalForSequenceClassification": ("llama_aliases", "LlamaBidirectionalForSequenceClassification"),
    "VoyageQwen3BidirectionalEmbedModel": ("llama_aliases", "VoyageQwen3BidirectionalEmbedModel"),
    "HyperCLOVAXForCausalLM": ("llama_variants", "HyperCLOVAXForCausalLM"),
    "Rnj1ForCausalLM": ("llama_variants", "Gemma3ForCausalLM"),
    "MellumForCausalLM": ("mixtral", "Qwen3MoeForCausalLM"),
    "SolarForCausalLM": ("llama", "LlamaForCausalLM"),
    "DeciLMForCausalLM": ("llama_aliases", "DeciLMForCausalLM"),
    "Gemma2ForCausalLM": ("gemma2", "Gemma2ForCausalLM"),
    "GPTNeoXForCausalLM": ("gpt_neox", "GPTNeoXForCausalLM"),
    "Starcoder2ForCausalLM": ("gpt_neox", "Starcoder2ForCausalLM"),
    "GPTJForCausalLM": ("gpt_neox", "GPTJForCausalLM"),
    "BloomForCausalLM": ("gpt_neox", "BloomForCausalLM"),
    "MptForCausalLM": ("gpt_neox", "MPTForCausalLM"),
    "MPTForCausalLM": ("gpt_neox", "MPTForCausalLM"),
    "FalconForCausalLM": ("gpt_neox", "FalconForCausalLM"),
    "RWForCausalLM": ("gpt_neox", "FalconForCausalLM"),
    "OPTForCausalLM": ("opt", "OPTForCausalLM"),
    "BioGptForCausalLM": ("opt", "BioGptForCausalLM"),
    "XGLMForCausalLM": ("opt", "XGLMForCausalLM"),
    "GPT2LMHeadModel": ("opt", "GPT2LMHeadModel"),
    "GPTBigCodeForCausalLM": ("opt", "GPTBigCodeForCausalLM"),
    "MixtralForCausalLM": ("mixtral", "MixtralForCausalLM"),
    "Qwen2MoeForCausalLM": ("mixtral", "Qwen2MoeForCausalLM"),
    "OlmoeForCausalLM": ("mixtral", "OlmoeForCausalLM"),
    "Qwen3MoeForCausalLM": ("mixtral", "Qwen3MoeForCausalLM"),
    "GptOssForCausalLM": ("gpt_oss", "GptOssForCausalLM"),
    "MambaForCausalLM": ("mamba", "MambaForCausalLM"),
    "FalconMambaForCausalLM": ("mamba", "MambaForCausalLM"),
    "Qwen3NextForCausalLM": ("qwen3_next", "Qwen3NextForCausalLM"),
    "Qwen3_5ForCausalLM": ("qwen3_next", "Qwen3_5ForCausalLM"),
    "Qwen3_5MoeForCausalLM": ("qwen3_next", "Qwen3_5MoeForCausalLM"),
    "NemotronHForCausalLM": ("nemotron_h", "NemotronHForCausalLM"),
    "NemotronHPuzzleForCausalLM": ("nemotron_h", "NemotronHForCausalLM"),
    "FalconH1ForCausalLM": ("falcon_h1", "FalconH1ForCausalLM"),
    "Lfm2ForCausalLM": ("lfm2", "Lfm2ForCausalLM"),
    "Lfm2MoeForCausalLM": ("lfm2", "Lfm2MoeForCausalLM"),
    "OlmoHybridForCausalLM": ("olmo_hybrid", "OlmoHybridForCausalLM"),
    "Mamba2ForCausalLM": ("mamba2", "Mamba2ForCausalLM"),
    "BambaForCausalLM": ("mamba2", "BambaForCausalLM"),
    "JambaForCausalLM": ("jamba", "JambaForCausalLM"),
    "BartForConditionalGeneration": ("bart", "BartForConditionalGeneration"),
    "BartModel": ("bart", "BartForConditionalGeneration"),
    "WhisperForConditionalGeneration": ("whisper", "WhisperForConditionalGeneration"),
    "Qwen2_5_VLForConditionalGeneration": ("qwen2_vl", "Qwen2_5_VLForConditionalGeneration"),
    "Qwen2VLForConditionalGeneration": ("qwen2_vl", "Qwen2VLForConditionalGeneration"),
    "Idefics3ForConditionalGeneration": ("idefics3", "Idefics3ForConditionalGeneration"),
    "SmolVLMForConditionalGeneration": ("idefics3", "SmolVLMForConditionalGeneration"),
    "Gemma3ForConditionalGeneration": ("idefics3", "Gemma3ForConditionalGeneration"),
    "Cohere2VisionForConditionalGeneration": ("idefics3", "Cohere2VisionForConditionalGeneration"),
    "AyaVisionForConditionalGeneration": ("idefics3", "AyaVisionForConditionalGeneration"),
    "InternVLForConditionalGeneration": ("internvl", "InternVLForConditionalGeneration"),
    "InternS1ForConditionalGeneration": ("internvl", "InternVLForConditionalGeneration"),
    "Qwen2AudioForConditionalGeneration": ("audio_lm", "Qwen2AudioForConditionalGeneration"),
    "UltravoxModel": ("audio_lm", "UltravoxModel"),
    "VoxtralForConditionalGeneration": ("audio_lm", "VoxtralForConditionalGeneration"),
    "Mistral3ForConditionalGeneration": ("mistral3", "Mistral3ForConditionalGeneration"),
    "LlavaNextForConditionalGeneration": ("llava_next", "LlavaNextForConditionalGeneration"),
    "Qwen3VLForConditionalGeneration": ("qwen3_vl", "Qwen3VLForConditionalGeneration"),
    "Qwen3VLMoeForConditionalGeneration": ("qwen3_vl", "Qwen3VLMoeForConditionalGeneration"),
    "LlavaOnevisionForConditionalGeneration": ("llava_next", "LlavaOnevisionForConditionalGeneration"),
    "PaliGemmaForConditionalGeneration": ("idefics3", "PaliGemmaForConditionalGeneration"),
    "Qwen3_5ForConditionalGeneration": ("qwen3_vl", "Qwen3_5ForConditionalGeneration"),
    "Qwen3_5MoeForConditionalGeneration": ("qwen3_vl", "Qwen3_5MoeForConditionalGeneration"),
    "LlavaNextVideoForConditionalGeneration": ("llava_next", "LlavaNextVideoForConditionalGeneration"),
    "Llama4ForConditionalGeneration": ("llama4_mm", "Llama4ForConditionalGeneration"),
    "LlavaForConditionalGeneration": ("llava", "LlavaForConditionalGeneration"),
    "ModernBertModel": ("modernbert", "ModernBertModel"),
    "ModernBertForMaskedLM": ("modernbert", "ModernBertModel"),
    "BertModel": ("bert", "BertEmbeddingModel"),
    "BertForMaskedLM": ("bert", "BertEmbeddingModel"),
    "BertForSequenceClassification": ("bert", "BertForSequenceClassification"),
    "RobertaModel": ("bert", "RobertaModel"),
    "XLMRobertaModel": ("bert", "RobertaModel"),
    "DeepseekV2ForCausalLM": ("deepseek", "DeepseekV3ForCausalLM"),
    "Llama4ForCausalLM": ("llama4", "Llama4ForCausalLM"),
    "Dots1ForCausalLM": ("deepseek", "Dots1ForCausalLM"),
    "Glm4MoeForCausalLM": ("deepseek", "Glm4MoeForCausalLM"),
    "Ernie4_5_MoeForCausalLM": ("deepseek", "Ernie4_5_MoeForCausalLM"),
    "DeepseekV3ForCausalLM": ("deepseek", "DeepseekV3ForCausalLM"),
}


_TEXT_BY_TYPE = {"llama": "LlamaForCausalLM", "mistral": "MistralForCausalLM", "qwen2": "Qwen2ForCausalLM", "qwen3": "Qwen3ForCausalLM",
                 "gemma": "GemmaForCausalLM", "gemma2": "Gemma2ForCausalLM", "gemma3_text": "Gemma3ForCausalLM", "phi3": "Phi3ForCausalLM",
                 "cohere2": "Cohere2ForCausalLM", "internlm2": "InternLM2ForCausalLM", "qwen3_moe": "Qwen3MoeForCausalLM",
                 "granite": "GraniteForCausalLM", "gpt_oss": "GptOssForCausalLM", "olmo2": "Olmo2ForCausalLM"}


def text_decoder_cls(cfg: dict, default: str = "LlamaForCausalLM") -> Type[nn.Module]:
    """The decoder class behind a composite (vision / audio + text) checkpoint, from its ``text_config.model_type``."""
    return resolve_model_cls(_TEXT_BY_TYPE.get(cfg.get("text_model_type") or "", default))


def register_model(arch: str, module: str, cls: str) -> None:
    """Plugin hook: ``module`` is a fully qualified module path or a name under vllm_b200.models."""
    _MODELS[arch] = (module, cls)


def supported_archs() -> list[str]:
    return sorted(_MODELS)


# bases whose causal-LM class is not called ``<X>ForCausalLM``
_IRREGULAR_BASE = {"GPT2ForCausalLM": "GPT2LMHeadModel", "Gemma3TextForCausalLM": "Gemma3ForCausalLM",
                   # pooling checkpoints routed to the Transformers back-end by name (``--model-impl transformers``)
                   "TransformersEmbeddingForCausalLM": "TransformersForCausalLM",
                   "TransformersMoEEmbeddingForCausalLM": "TransformersForCausalLM",
                   "TransformersMoEForCausalLM": "TransformersForCausalLM"}
_REWARD = {"Qwen2ForRewardModel": ("Qwen2ForCausalLM", dict(n_out=1, mlp=True)),
           "Qwen2ForProcessRewardModel": ("Qwen2ForCausalLM", dict(n_out=2, mlp=True)),
           "InternLM2ForRewardModel": ("InternLM2ForCausalLM", dict(n_out=1, mlp=False, head_name="v_head"))}


def _adapter_for(arch: str):
    """``XForSequenceClassification`` / ``XForRewardModel`` / bare ``XModel`` of a registered ``XForCausalLM``."""
    if arch in _REWARD:
        return _REWARD[arch][0], "reward"
    for suffix, kind in (("ForSequenceClassification", "cls"), ("ForRewardModel", "cls"), ("Model", "embed")):
        if arch.endswith(suffix):
            base = arch[: -len(suffix)] + "ForCausalLM"
            base = _IRREGULAR_BASE.get(base, base)
            if base in _MODELS:
                return base, kind
    return None


def resolve_model_cls(arch: str, impl: str = "auto") -> Type[nn.Module]:
    """``impl``: ``auto`` — native class, else the Transformers back-end for causal LMs transformers implements;
    ``vllm_b200`` (alias ``vllm``) — native classes only; ``transformers`` — always the back-end (reference: ``--model-impl``)."""
    if impl == "transformers":
        from vllm_b200.models import adapters
        from vllm_b200.models.transformers_backend import TransformersForCausalLM
        if arch.endswith("ForSequenceClassification") or arch.endswith("ForRewardModel"):
            return adapters.as_seq_cls_model(TransformersForCausalLM)
        if arch.endswith("Model") and not arch.endswith("LMHeadModel"):
            return adapters.as_embedding_model(TransformersForCausalLM)
        return TransformersForCausalLM
    if impl in ("vllm", "vllm_b200") and arch not in _MODELS and not _adapter_for(arch):
        raise ValueError(f"architecture {arch!r} has no native implementation (--model-impl {impl}); supported: {supported_archs()}")
    if arch not in _MODELS and _adapter_for(arch):
        from vllm_b200.models import adapters
        base, kind = _adapter_for(arch)
        if kind == "reward":
            return adapters.as_reward_model(resolve_model_cls(base), **_REWARD[arch][1])
        fn = adapters.as_seq_cls_model if kind == "cls" else adapters.as_embedding_model
        return fn(resolve_model_cls(base))
    if arch not in _MODELS:
        # the long tail: any decoder-only HF architecture runs through the Transformers back-end
        try:
            import transformers
            if arch.endswith("ForCausalLM") and hasattr(transformers, arch):
                from vllm_b200.models.transformers_backend import TransformersForCausalLM
                return TransformersForCausalLM
        except ImportError:
            pass
        raise ValueError(f"architecture {arch!r} is not supported; supported: {supported_archs()}")
    mod, cls = _MODELS[arch]
    if "." not in mod:
        mod = f"vllm_b200.models.{mod}"
    return getattr(importlib.import_module(mod), cls)
