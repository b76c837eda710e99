"""ModelConfig: architecture description taken from a HuggingFace ``config.json``
or from a built-in preset (there is no network in the build environment, so the
benchmark architectures are baked in).

Reference: vllm/config/model.py (``ModelConfig``), transformers_utils/config.py.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Any, Optional

import torch

_DTYPES = {
    "auto": None, "bfloat16": torch.bfloat16, "bf16": torch.bfloat16, "float16": torch.float16,
    "half": torch.float16, "fp16": torch.float16, "float32": torch.float32, "float": torch.float32,
}

# HF-style config dicts for the architectures named in BASELINE.json plus tiny
# variants used by CPU tests.
PRESETS: dict[str, dict[str, Any]] = {
    "opt-125m": dict(
        architectures=["OPTForCausalLM"], model_type="opt", vocab_size=50272, hidden_size=768,
        num_hidden_layers=12, num_attention_heads=12, ffn_dim=3072, max_position_embeddings=2048,
        word_embed_proj_dim=768, do_layer_norm_before=True, activation_function="relu",
        bos_token_id=2, eos_token_id=2, pad_token_id=1, torch_dtype="float16", enable_bias=True,
        layer_norm_elementwise_affine=True, tie_word_embeddings=True),
    "opt-tiny": dict(
        architectures=["OPTForCausalLM"], model_type="opt", vocab_size=512, hidden_size=64,
        num_hidden_layers=2, num_attention_heads=4, ffn_dim=256, max_position_embeddings=512,
        word_embed_proj_dim=64, do_layer_norm_before=True, activation_function="relu",
        bos_token_id=2, eos_token_id=2, pad_token_id=1, torch_dtype="float32", enable_bias=True,
        layer_norm_elementwise_affine=True, tie_word_embeddings=True),
    "llama-3-8b": dict(
        architectures=["LlamaForCausalLM"], model_type="llama", vocab_size=128256, hidden_size=4096,
        intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8,
        head_dim=128, max_position_embeddings=8192, rms_norm_eps=1e-5, rope_theta=500000.0,
        hidden_act="silu", bos_token_id=128000, eos_token_id=128001, torch_dtype="bfloat16",
        tie_word_embeddings=False, attention_bias=False, mlp_bias=False),
    # one rank's share of Llama-3-8B under TP=8 (4 of 32 heads, 1 of 8 KV heads, 1/8 of the MLP and of the vocabulary):
    # lets a single GPU run — and ncu profile — exactly the kernel shapes of a TP8 decode step, without the collectives
    "llama-3-8b-tp8-shard": dict(
        architectures=["LlamaForCausalLM"], model_type="llama", vocab_size=16032, hidden_size=4096,
        intermediate_size=1792, num_hidden_layers=32, num_attention_heads=4, num_key_value_heads=1,
        head_dim=128, max_position_embeddings=8192, rms_norm_eps=1e-5, rope_theta=500000.0,
        hidden_act="silu", bos_token_id=0, eos_token_id=1, torch_dtype="bfloat16",
        tie_word_embeddings=False, attention_bias=False, mlp_bias=False),
    "llama-3-70b": dict(
        architectures=["LlamaForCausalLM"], model_type="llama", vocab_size=128256, hidden_size=8192,
        intermediate_size=28672, num_hidden_layers=80, num_attention_heads=64, num_key_value_heads=8,
        head_dim=128, max_position_embeddings=8192, rms_norm_eps=1e-5, rope_theta=500000.0,
        hidden_act="silu", bos_token_id=128000, eos_token_id=128001, torch_dtype="bfloat16",
        tie_word_embeddings=False, attention_bias=False, mlp_bias=False),
    "llama-tiny": dict(
        architectures=["LlamaForCausalLM"], model_type="llama", vocab_size=512, hidden_size=128,
        intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
        head_dim=32, max_position_embeddings=1024, rms_norm_eps=1e-5, rope_theta=10000.0,
        hidden_act="silu", bos_token_id=1, eos_token_id=2, torch_dtype="float32",
        tie_word_embeddings=False),
    # GPU-test sized llama: head_dim 128 so the sm_100a attention kernels apply
    "llama-small": dict(
        architectures=["LlamaForCausalLM"], model_type="llama", vocab_size=4096, hidden_size=1024,
        intermediate_size=2816, num_hidden_layers=4, num_attention_heads=8, num_key_value_heads=2,
        head_dim=128, max_position_embeddings=4096, rms_norm_eps=1e-5, rope_theta=500000.0,
        hidden_act="silu", bos_token_id=1, eos_token_id=2, torch_dtype="bfloat16",
        tie_word_embeddings=False),
    "qwen2.5-7b": dict(
        architectures=["Qwen2ForCausalLM"], model_type="qwen2", vocab_size=152064, hidden_size=3584,
        intermediate_size=18944, num_hidden_layers=28, num_attention_heads=28, num_key_value_heads=4,
        max_position_embeddings=32768, rms_norm_eps=1e-6, rope_theta=1000000.0, hidden_act="silu",
        bos_token_id=151643, eos_token_id=151645, torch_dtype="bfloat16", tie_word_embeddings=False,
        attention_bias=True),
    "qwen3-8b": dict(
        architectures=["Qwen3ForCausalLM"], model_type="qwen3", vocab_size=151936, hidden_size=4096,
        intermediate_size=12288, num_hidden_layers=36, num_attention_heads=32, num_key_value_heads=8,
        head_dim=128, max_position_embeddings=40960, rms_norm_eps=1e-6, rope_theta=1000000.0,
        hidden_act="silu", bos_token_id=151643, eos_token_id=151645, torch_dtype="bfloat16",
        tie_word_embeddings=False, attention_bias=False),
    "mistral-7b": dict(
        architectures=["MistralForCausalLM"], model_type="mistral", vocab_size=32000, hidden_size=4096,
        intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8,
        max_position_embeddings=32768, rms_norm_eps=1e-5, rope_theta=10000.0, hidden_act="silu",
        sliding_window=4096, bos_token_id=1, eos_token_id=2, torch_dtype="bfloat16",
        tie_word_embeddings=False),
    "gemma-2b": dict(
        architectures=["GemmaForCausalLM"], model_type="gemma", vocab_size=256000, hidden_size=2048,
        intermediate_size=16384, num_hidden_layers=18, num_attention_heads=8, num_key_value_heads=1,
        head_dim=256, max_position_embeddings=8192, rms_norm_eps=1e-6, rope_theta=10000.0,
        hidden_act="gelu_pytorch_tanh", bos_token_id=2, eos_token_id=1, torch_dtype="bfloat16",
        tie_word_embeddings=True),
    "gpt2": dict(
        architectures=["GPT2LMHeadModel"], model_type="gpt2", vocab_size=50257, n_embd=768, n_layer=12,
        n_head=12, n_positions=1024, n_inner=None, activation_function="gelu_new",
        layer_norm_epsilon=1e-5, bos_token_id=50256, eos_token_id=50256, torch_dtype="float32",
        tie_word_embeddings=True),
    "gpt2-tiny": dict(
        architectures=["GPT2LMHeadModel"], model_type="gpt2", vocab_size=512, n_embd=64, n_layer=2,
        n_head=4, n_positions=512, n_inner=None, activation_function="gelu_new",
        layer_norm_epsilon=1e-5, bos_token_id=1, eos_token_id=2, torch_dtype="float32",
        tie_word_embeddings=True),
    "mixtral-8x7b": dict(
        architectures=["MixtralForCausalLM"], model_type="mixtral", vocab_size=32000, hidden_size=4096,
        intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8,
        max_position_embeddings=32768, rms_norm_eps=1e-5, rope_theta=1000000.0, hidden_act="silu",
        num_local_experts=8, num_experts_per_tok=2, bos_token_id=1, eos_token_id=2,
        torch_dtype="bfloat16", tie_word_embeddings=False),
    "mixtral-tiny": dict(
        architectures=["MixtralForCausalLM"], model_type="mixtral", vocab_size=512, hidden_size=128,
        intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
        head_dim=32, max_position_embeddings=1024, rms_norm_eps=1e-5, rope_theta=10000.0,
        hidden_act="silu", num_local_experts=4, num_experts_per_tok=2, bos_token_id=1, eos_token_id=2,
        torch_dtype="float32", tie_word_embeddings=False),
    "mixtral-small": dict(
        architectures=["MixtralForCausalLM"], model_type="mixtral", vocab_size=4096, hidden_size=1024,
        intermediate_size=2048, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=2,
        head_dim=128, max_position_embeddings=4096, rms_norm_eps=1e-5, rope_theta=10000.0,
        hidden_act="silu", num_local_experts=8, num_experts_per_tok=2, bos_token_id=1, eos_token_id=2,
        torch_dtype="bfloat16", tie_word_embeddings=False),
    "llava-tiny": dict(
        architectures=["LlavaForConditionalGeneration"], model_type="llava", image_token_index=511,
        vision_feature_layer=-2, vision_feature_select_strategy="default", projector_hidden_act="gelu",
        vision_config=dict(model_type="clip_vision_model", hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                           num_attention_heads=2, image_size=28, patch_size=14, projection_dim=32,
                           hidden_act="quick_gelu", layer_norm_eps=1e-5),
        # text stack (flattened text_config)
        vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=2, max_position_embeddings=256, rms_norm_eps=1e-6, rope_theta=10000.0,
        hidden_act="silu", bos_token_id=1, eos_token_id=2, torch_dtype="float32", tie_word_embeddings=False),
    "smolvlm-tiny": dict(
        architectures=["SmolVLMForConditionalGeneration"], model_type="smolvlm", image_token_id=511, scale_factor=2,
        vision_config=dict(model_type="smolvlm_vision", hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                           num_attention_heads=2, image_size=32, patch_size=8, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6),
        vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=2, max_position_embeddings=256, rms_norm_eps=1e-6, rope_theta=10000.0,
        hidden_act="silu", bos_token_id=1, eos_token_id=2, torch_dtype="float32", tie_word_embeddings=False),
    "gemma3mm-tiny": dict(
        architectures=["Gemma3ForConditionalGeneration"], model_type="gemma3", image_token_index=511, mm_tokens_per_image=4,
        vision_config=dict(model_type="siglip_vision_model", hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                           num_attention_heads=2, image_size=32, patch_size=8, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6,
                           vision_use_head=False),
        vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
        max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", head_dim=16,
        rms_norm_eps=1e-6, hidden_activation="gelu_pytorch_tanh", query_pre_attn_scalar=16, sliding_window=16,
        layer_types=["sliding_attention", "full_attention"],
        rope_parameters=dict(sliding_attention=dict(rope_type="default", rope_theta=10000.0),
                             full_attention=dict(rope_type="default", rope_theta=1000000.0)),
        tie_word_embeddings=True),
    "cohere2vision-tiny": dict(
        architectures=["Cohere2VisionForConditionalGeneration"], model_type="cohere2_vision", image_token_id=511, downsample_factor=2,
        alignment_intermediate_size=96,
        vision_config=dict(model_type="siglip_vision_model", hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                           num_attention_heads=2, image_size=32, patch_size=8, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6,
                           vision_use_head=False),
        vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2,
        max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", rope_theta=10000.0,
        hidden_act="silu", head_dim=16, layer_norm_eps=1e-5, logit_scale=0.25,
        layer_types=["sliding_attention", "sliding_attention", "sliding_attention", "full_attention"], sliding_window=16,
        tie_word_embeddings=True),
    "internvl-tiny": dict(
        architectures=["InternVLForConditionalGeneration"], model_type="internvl", image_token_id=511, downsample_ratio=0.5,
        vision_feature_layer=-1, vision_feature_select_strategy="default", projector_hidden_act="gelu", text_model_type="qwen2",
        vision_config=dict(model_type="internvl_vision", hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                           num_attention_heads=2, image_size=[32, 32], patch_size=[8, 8], use_qk_norm=True, norm_type="layer_norm",
                           attention_bias=True, layer_norm_eps=1e-6, hidden_act="gelu", use_mean_pooling=True,
                           use_absolute_position_embeddings=True, layer_scale_init_value=0.5),
        vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=2, max_position_embeddings=256, rms_norm_eps=1e-6, rope_theta=10000.0,
        hidden_act="silu", bos_token_id=1, eos_token_id=2, torch_dtype="float32", tie_word_embeddings=False),
    "qwen2audio-tiny": dict(
        architectures=["Qwen2AudioForConditionalGeneration"], model_type="qwen2_audio", audio_token_index=511, text_model_type="qwen2",
        audio_config=dict(model_type="qwen2_audio_encoder", num_mel_bins=16, d_model=32, encoder_layers=2, encoder_attention_heads=2,
                          encoder_ffn_dim=64, max_source_positions=24, activation_function="gelu", scale_embedding=False),
        vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=2, max_position_embeddings=256, rms_norm_eps=1e-6, rope_theta=10000.0,
        hidden_act="silu", bos_token_id=1, eos_token_id=2, torch_dtype="float32", tie_word_embeddings=False),
    "ultravox-tiny": dict(
        architectures=["UltravoxModel"], model_type="ultravox", audio_token_index=511, text_model_type="llama", stack_factor=4,
        projector_act="swiglu", projector_ln_mid=True, projector_hidden_size=96,
        audio_config=dict(model_type="whisper", num_mel_bins=16, d_model=32, encoder_layers=2, encoder_attention_heads=2,
                          encoder_ffn_dim=64, max_source_positions=24),
        vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=2, max_position_embeddings=256, rms_norm_eps=1e-6, rope_theta=10000.0,
        hidden_act="silu", bos_token_id=1, eos_token_id=2, torch_dtype="float32", tie_word_embeddings=False),
    "voxtral-tiny": dict(
        architectures=["VoxtralForConditionalGeneration"], model_type="voxtral", audio_token_id=511, text_model_type="llama",
        projector_hidden_act="gelu",
        audio_config=dict(model_type="voxtral_encoder", vocab_size=512, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                          num_attention_heads=2, num_mel_bins=16, max_source_positions=24, head_dim=16),
        vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=2, head_dim=16, max_position_embeddings=256, rms_norm_eps=1e-6, rope_theta=10000.0,
        hidden_act="silu", bos_token_id=1, eos_token_id=2, torch_dtype="float32", tie_word_embeddings=False),
    "mistral3-tiny": dict(
        architectures=["Mistral3ForConditionalGeneration"], model_type="mistral3", image_token_index=10, spatial_merge_size=2,
        multimodal_projector_bias=False, projector_hidden_act="gelu", vision_feature_layer=-1, text_model_type="mistral",
        vision_config=dict(model_type="pixtral", hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2,
                           head_dim=16, image_size=32, patch_size=4, hidden_act="silu", rope_theta=10000.0),
        vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=2, head_dim=16, max_position_embeddings=256, rms_norm_eps=1e-5, rope_theta=10000.0, sliding_window=None,
        hidden_act="silu", bos_token_id=1, eos_token_id=2, torch_dtype="float32", tie_word_embeddings=False),
    "llavaov-tiny": dict(
        architectures=["LlavaOnevisionForConditionalGeneration"], model_type="llava_onevision", image_token_index=511,
        vision_feature_layer=-1, vision_feature_select_strategy="full", projector_hidden_act="gelu", vision_aspect_ratio="anyres_max_2",
        image_grid_pinpoints=[[32, 32], [32, 64], [64, 32], [64, 64], [32, 96], [96, 32]], multimodal_projector_bias=True,
        text_model_type="qwen2",
        vision_config=dict(model_type="siglip_vision_model", hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                           num_attention_heads=2, image_size=32, patch_size=8, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6,
                           vision_use_head=False),
        vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=2, max_position_embeddings=512, rms_norm_eps=1e-6, rope_theta=10000.0,
        hidden_act="silu", bos_token_id=1, eos_token_id=2, torch_dtype="float32", tie_word_embeddings=False),
    "paligemma-tiny": dict(
        architectures=["PaliGemmaForConditionalGeneration"], model_type="paligemma", image_token_index=511, text_model_type="gemma",
        projection_dim=64,
        vision_config=dict(model_type="siglip_vision_model", hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                           num_attention_heads=2, image_size=32, patch_size=16, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6,
                           vision_use_head=False, projection_dim=64),
        vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=1,
        head_dim=16, max_position_embeddings=256, rms_norm_eps=1e-6, rope_theta=10000.0, hidden_act="gelu_pytorch_tanh",
        hidden_activation="gelu_pytorch_tanh", bos_token_id=2, eos_token_id=1, pad_token_id=0, torch_dtype="float32",
        tie_word_embeddings=True),
    "mamba-tiny": dict(
        architectures=["MambaForCausalLM"], model_type="mamba", vocab_size=512, hidden_size=64, state_size=8,
        num_hidden_layers=2, expand=2, conv_kernel=4, time_step_rank=4, use_bias=False, use_conv_bias=True,
        layer_norm_epsilon=1e-5, intermediate_size=128, bos_token_id=0, eos_token_id=0, pad_token_id=0,
        torch_dtype="float32", tie_word_embeddings=True, max_position_embeddings=1024),
    "biogpt-tiny": dict(
        architectures=["BioGptForCausalLM"], model_type="biogpt", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=2, num_attention_heads=4, max_position_embeddings=256, hidden_act="gelu", layer_norm_eps=1e-5,
        scale_embedding=True, bos_token_id=0, eos_token_id=2, pad_token_id=1, torch_dtype="float32",
        tie_word_embeddings=True, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, activation_dropout=0.0),
    "xglm-tiny": dict(
        architectures=["XGLMForCausalLM"], model_type="xglm", vocab_size=512, d_model=64, ffn_dim=128, num_layers=2,
        attention_heads=4, max_position_embeddings=256, activation_function="gelu", scale_embedding=True,
        bos_token_id=0, eos_token_id=2, pad_token_id=1, decoder_start_token_id=2, torch_dtype="float32",
        tie_word_embeddings=True, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, layerdrop=0.0),
    "hunyuan-moe-tiny": dict(
        architectures=["HunYuanMoEV1ForCausalLM"], model_type="hunyuan_v1_moe", vocab_size=512, hidden_size=64,
        intermediate_size=96, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=16, num_experts=4,
        moe_topk=2, max_position_embeddings=256, hidden_act="silu", rms_norm_eps=1e-5, rope_theta=10000.0, use_qk_norm=True,
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "ernie45moe-tiny": dict(
        architectures=["Ernie4_5_MoeForCausalLM"], model_type="ernie4_5_moe", vocab_size=512, hidden_size=64,
        intermediate_size=128, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2, moe_num_experts=4, moe_k=2,
        moe_intermediate_size=48, moe_num_shared_experts=1, moe_layer_start_index=1, moe_layer_interval=1,
        moe_layer_end_index=2, max_position_embeddings=256, rms_norm_eps=1e-5, rope_theta=10000.0, hidden_act="silu",
        use_bias=False, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=True),
    "vaultgemma-tiny": dict(
        architectures=["VaultGemmaForCausalLM"], model_type="vaultgemma", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=256,
        sliding_window=16, layer_types=["sliding_attention", "full_attention", "sliding_attention", "full_attention"],
        query_pre_attn_scalar=16, attn_logit_softcapping=50.0, final_logit_softcapping=30.0,
        hidden_activation="gelu_pytorch_tanh", rms_norm_eps=1e-6, rope_theta=10000.0, bos_token_id=2, eos_token_id=1,
        pad_token_id=0, torch_dtype="float32", tie_word_embeddings=True),
    "falconmamba-tiny": dict(
        architectures=["FalconMambaForCausalLM"], model_type="falcon_mamba", vocab_size=512, hidden_size=64, state_size=8,
        num_hidden_layers=2, expand=2, conv_kernel=4, time_step_rank=4, use_bias=False, use_conv_bias=True,
        layer_norm_epsilon=1e-5, mixer_rms_eps=1e-6, intermediate_size=128, bos_token_id=0, eos_token_id=0, pad_token_id=0,
        torch_dtype="float32", tie_word_embeddings=True, max_position_embeddings=1024),
    "phi3-tiny": dict(
        architectures=["Phi3ForCausalLM"], model_type="phi3", vocab_size=512, hidden_size=128, intermediate_size=256,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=1024,
        original_max_position_embeddings=1024, rms_norm_eps=1e-5, rope_theta=10000.0, hidden_act="silu",
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "gemma2-tiny": dict(
        architectures=["Gemma2ForCausalLM"], model_type="gemma2", vocab_size=512, hidden_size=128, intermediate_size=256,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=32, max_position_embeddings=1024,
        rms_norm_eps=1e-6, rope_theta=10000.0, hidden_activation="gelu_pytorch_tanh", hidden_act="gelu_pytorch_tanh",
        query_pre_attn_scalar=32, attn_logit_softcapping=50.0, final_logit_softcapping=30.0, sliding_window=16,
        bos_token_id=2, eos_token_id=1, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=True),
    "gptneox-tiny": dict(
        architectures=["GPTNeoXForCausalLM"], model_type="gpt_neox", vocab_size=512, hidden_size=64, intermediate_size=256,
        num_hidden_layers=2, num_attention_heads=4, max_position_embeddings=512, rotary_pct=0.25, rotary_emb_base=10000,
        layer_norm_eps=1e-5, hidden_act="gelu", use_parallel_residual=True, bos_token_id=0, eos_token_id=0,
        torch_dtype="float32", tie_word_embeddings=False),
    "starcoder2-tiny": dict(
        architectures=["Starcoder2ForCausalLM"], model_type="starcoder2", vocab_size=512, hidden_size=64,
        intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
        max_position_embeddings=512, rope_theta=10000.0, norm_epsilon=1e-5, hidden_act="gelu_pytorch_tanh", use_bias=True,
        sliding_window=None, bos_token_id=0, eos_token_id=0, torch_dtype="float32", tie_word_embeddings=True),
    "gptj-tiny": dict(
        architectures=["GPTJForCausalLM"], model_type="gptj", vocab_size=512, n_embd=64, n_layer=2, n_head=4, n_inner=256,
        n_positions=512, rotary_dim=8, layer_norm_epsilon=1e-5, activation_function="gelu_new", bos_token_id=0,
        eos_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "bloom-tiny": dict(
        architectures=["BloomForCausalLM"], model_type="bloom", vocab_size=512, hidden_size=64, n_layer=2, n_head=4,
        num_hidden_layers=2, num_attention_heads=4, layer_norm_epsilon=1e-5, bos_token_id=1, eos_token_id=2,
        torch_dtype="float32", tie_word_embeddings=True, max_position_embeddings=512),
    "mpt-tiny": dict(
        architectures=["MptForCausalLM"], model_type="mpt", vocab_size=512, d_model=64, n_heads=4, n_layers=2,
        hidden_size=64, num_hidden_layers=2, num_attention_heads=4, expansion_ratio=4, max_seq_len=512,
        max_position_embeddings=512, no_bias=True, layer_norm_epsilon=1e-5,
        attn_config=dict(alibi=True, alibi_bias_max=8, attn_type="multihead_attention", attn_impl="torch"),
        torch_dtype="float32", tie_word_embeddings=True),
    "falcon-rw-tiny": dict(
        architectures=["FalconForCausalLM"], model_type="falcon", vocab_size=512, hidden_size=64, num_hidden_layers=2,
        num_attention_heads=4, alibi=True, bias=True, multi_query=False, parallel_attn=False,
        new_decoder_architecture=False, layer_norm_epsilon=1e-5, bos_token_id=1, eos_token_id=2,
        max_position_embeddings=512, torch_dtype="float32", tie_word_embeddings=True),
    "falcon-7b-tiny": dict(
        architectures=["FalconForCausalLM"], model_type="falcon", vocab_size=512, hidden_size=64, num_hidden_layers=2,
        num_attention_heads=4, alibi=False, bias=False, multi_query=True, parallel_attn=True,
        new_decoder_architecture=False, layer_norm_epsilon=1e-5, bos_token_id=1, eos_token_id=2, rope_theta=10000.0,
        max_position_embeddings=512, torch_dtype="float32", tie_word_embeddings=True),
    "falcon-40b-tiny": dict(
        architectures=["FalconForCausalLM"], model_type="falcon", vocab_size=512, hidden_size=64, num_hidden_layers=2,
        num_attention_heads=4, num_kv_heads=2, alibi=False, bias=False, multi_query=True, parallel_attn=True,
        new_decoder_architecture=True, layer_norm_epsilon=1e-5, bos_token_id=1, eos_token_id=2, rope_theta=10000.0,
        max_position_embeddings=512, torch_dtype="float32", tie_word_embeddings=True),
    "gptoss-tiny": dict(
        architectures=["GptOssForCausalLM"], model_type="gpt_oss", vocab_size=512, hidden_size=64, intermediate_size=64,
        num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16, num_local_experts=4,
        num_experts_per_tok=2, sliding_window=16, max_position_embeddings=512, rms_norm_eps=1e-5, attention_bias=True,
        layer_types=["sliding_attention", "full_attention", "sliding_attention", "full_attention"],
        rope_parameters=dict(rope_type="yarn", factor=4.0, beta_fast=32.0, beta_slow=1.0, truncate=False,
                             original_max_position_embeddings=128, rope_theta=10000.0),
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "mamba2-tiny": dict(
        architectures=["Mamba2ForCausalLM"], model_type="mamba2", vocab_size=512, hidden_size=64, num_hidden_layers=2,
        num_heads=8, head_dim=16, expand=2, state_size=16, n_groups=2, conv_kernel=4, chunk_size=16, use_bias=False,
        use_conv_bias=True, layer_norm_epsilon=1e-5, rms_norm=True, time_step_limit=[0.0, float("inf")],
        bos_token_id=0, eos_token_id=0, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False,
        num_attention_heads=8, max_position_embeddings=512),
    "bamba-tiny": dict(
        architectures=["BambaForCausalLM"], model_type="bamba", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, attn_layer_indices=[1, 3],
        mamba_n_heads=8, mamba_d_head=16, mamba_expand=2, mamba_d_state=16, mamba_n_groups=2, mamba_d_conv=4,
        mamba_chunk_size=16, mamba_conv_bias=True, mamba_proj_bias=False, rms_norm_eps=1e-5, hidden_act="silu",
        max_position_embeddings=512, rope_theta=10000.0, partial_rotary_factor=1.0, attention_bias=False,
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "nemotronh-tiny": dict(
        architectures=["NemotronHForCausalLM"], model_type="nemotron_h", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=4, layers_block_type=["mamba", "moe", "attention", "moe"], num_attention_heads=4, num_key_value_heads=2,
        head_dim=16, mamba_num_heads=4, mamba_head_dim=16, ssm_state_size=16, n_groups=2, conv_kernel=4, chunk_size=32,
        use_conv_bias=True, use_bias=False, mlp_bias=False, mlp_hidden_act="relu2", mamba_hidden_act="silu",
        n_routed_experts=8, num_experts_per_tok=2, moe_intermediate_size=32, moe_shared_expert_intermediate_size=48,
        moe_latent_size=None, n_group=2, topk_group=1, norm_topk_prob=True, routed_scaling_factor=1.5, layer_norm_epsilon=1e-5,
        max_position_embeddings=256, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32",
        tie_word_embeddings=False),
    "nemotronh-latent-tiny": dict(
        architectures=["NemotronHForCausalLM"], model_type="nemotron_h", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=3, layers_block_type=["mamba", "attention", "moe"], num_attention_heads=4, num_key_value_heads=4,
        head_dim=16, mamba_num_heads=8, mamba_head_dim=8, ssm_state_size=16, n_groups=1, conv_kernel=4, chunk_size=32,
        use_conv_bias=True, use_bias=True, mlp_bias=True, mlp_hidden_act="relu2", mamba_hidden_act="silu",
        n_routed_experts=4, num_experts_per_tok=2, moe_intermediate_size=32, moe_shared_expert_intermediate_size=48,
        moe_latent_size=32, n_group=1, topk_group=1, norm_topk_prob=False, routed_scaling_factor=1.0, layer_norm_epsilon=1e-5,
        max_position_embeddings=256, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32",
        tie_word_embeddings=False),
    "falconh1-tiny": dict(
        architectures=["FalconH1ForCausalLM"], model_type="falcon_h1", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=256,
        mamba_d_ssm=64, mamba_n_heads=4, mamba_d_head=16, mamba_n_groups=2, mamba_d_state=16, mamba_d_conv=4, mamba_chunk_size=32,
        mamba_conv_bias=True, mamba_proj_bias=False, mamba_norm_before_gate=True, mamba_rms_norm=True, projectors_bias=False,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0), lm_head_multiplier=0.7, embedding_multiplier=3.0,
        mlp_multipliers=[0.9, 1.2], key_multiplier=0.8, attention_out_multiplier=0.6, attention_in_multiplier=1.3,
        ssm_multipliers=[0.5, 1.5, 0.7, 1.1, 0.9], ssm_in_multiplier=1.4, ssm_out_multiplier=0.5, attention_bias=False,
        mlp_bias=False, hidden_act="silu", rms_norm_eps=1e-5, bos_token_id=1, eos_token_id=2, pad_token_id=0,
        torch_dtype="float32", tie_word_embeddings=False),
    "falconh1-nonorm-tiny": dict(
        architectures=["FalconH1ForCausalLM"], model_type="falcon_h1", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=256,
        mamba_d_ssm=64, mamba_n_heads=4, mamba_d_head=16, mamba_n_groups=1, mamba_d_state=16, mamba_d_conv=4, mamba_chunk_size=32,
        mamba_conv_bias=True, mamba_proj_bias=True, mamba_rms_norm=False, projectors_bias=True,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0), lm_head_multiplier=1.0, embedding_multiplier=1.0,
        mlp_multipliers=[1.1, 0.8], key_multiplier=1.2, attention_out_multiplier=0.9, attention_in_multiplier=1.0,
        ssm_multipliers=[1.0, 0.8, 1.2, 1.0, 1.1], ssm_in_multiplier=1.0, ssm_out_multiplier=1.3, attention_bias=True,
        mlp_bias=True, hidden_act="silu", rms_norm_eps=1e-5, bos_token_id=1, eos_token_id=2, pad_token_id=0,
        torch_dtype="float32", tie_word_embeddings=True),
    "lfm2-tiny": dict(
        architectures=["Lfm2ForCausalLM"], model_type="lfm2", vocab_size=512, hidden_size=64, intermediate_size=192,
        num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=256, norm_eps=1e-5,
        conv_bias=False, conv_L_cache=3, block_multiple_of=32, block_ffn_dim_multiplier=1.0, block_auto_adjust_ff_dim=True,
        layer_types=["conv", "conv", "full_attention", "conv"], rope_parameters=dict(rope_type="default", rope_theta=10000.0),
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=True),
    "lfm2moe-tiny": dict(
        architectures=["Lfm2MoeForCausalLM"], model_type="lfm2_moe", vocab_size=512, hidden_size=64, intermediate_size=128,
        moe_intermediate_size=32, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=256,
        norm_eps=1e-5, conv_bias=True, conv_L_cache=4, num_dense_layers=1, num_experts_per_tok=2, num_experts=8, use_expert_bias=True,
        routed_scaling_factor=1.5, norm_topk_prob=True, layer_types=["conv", "full_attention", "conv", "full_attention"],
        rope_parameters=dict(rope_type="default", rope_theta=10000.0), bos_token_id=1, eos_token_id=2, pad_token_id=0,
        torch_dtype="float32", tie_word_embeddings=True),
    "olmohybrid-tiny": dict(
        architectures=["OlmoHybridForCausalLM"], model_type="olmo_hybrid", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=256,
        layer_types=["linear_attention", "full_attention", "linear_attention", "full_attention"], linear_num_key_heads=2,
        linear_num_value_heads=4, linear_key_head_dim=16, linear_value_head_dim=16, linear_conv_kernel_dim=4,
        linear_allow_neg_eigval=True, rms_norm_eps=1e-6, hidden_act="silu", rope_parameters=dict(rope_type="default", rope_theta=10000.0),
        pad_token_id=0, eos_token_id=2, torch_dtype="float32", tie_word_embeddings=False),
    "qwen35-tiny": dict(
        architectures=["Qwen3_5ForCausalLM"], model_type="qwen3_5_text", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16, linear_num_value_heads=4,
        linear_num_key_heads=2, linear_key_head_dim=16, linear_value_head_dim=16, linear_conv_kernel_dim=4,
        layer_types=["linear_attention", "full_attention", "linear_attention", "full_attention"], rms_norm_eps=1e-6,
        hidden_act="silu", max_position_embeddings=512, attention_bias=False,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0, partial_rotary_factor=0.25), bos_token_id=1,
        eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "qwen35moe-tiny": dict(
        architectures=["Qwen3_5MoeForCausalLM"], model_type="qwen3_5_moe_text", vocab_size=512, hidden_size=64,
        num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16, linear_num_value_heads=4,
        linear_num_key_heads=2, linear_key_head_dim=16, linear_value_head_dim=16, linear_conv_kernel_dim=4,
        layer_types=["linear_attention", "full_attention", "linear_attention", "full_attention"],
        num_experts=4, num_experts_per_tok=2, moe_intermediate_size=32, shared_expert_intermediate_size=32, rms_norm_eps=1e-6,
        hidden_act="silu", max_position_embeddings=512, attention_bias=False,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0, partial_rotary_factor=0.25), bos_token_id=1,
        eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "qwen3next-tiny": dict(
        architectures=["Qwen3NextForCausalLM"], model_type="qwen3_next", vocab_size=512, hidden_size=64,
        intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16,
        linear_num_value_heads=4, linear_num_key_heads=2, linear_key_head_dim=16, linear_value_head_dim=16,
        linear_conv_kernel_dim=4, full_attention_interval=2,
        layer_types=["linear_attention", "full_attention", "linear_attention", "full_attention"],
        num_experts=4, num_experts_per_tok=2, moe_intermediate_size=32, shared_expert_intermediate_size=32,
        decoder_sparse_step=1, norm_topk_prob=True, mlp_only_layers=[], rms_norm_eps=1e-6, hidden_act="silu",
        max_position_embeddings=512, attention_bias=False,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0, partial_rotary_factor=0.25),
        partial_rotary_factor=0.25, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32",
        tie_word_embeddings=False),
    "qwen2vl-tiny": dict(
        architectures=["Qwen2VLForConditionalGeneration"], model_type="qwen2_vl", vocab_size=512, hidden_size=64,
        intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
        max_position_embeddings=512, rms_norm_eps=1e-6, hidden_act="silu", rope_theta=10000.0, attention_bias=True,
        rope_scaling=dict(rope_type="mrope", mrope_section=[2, 3, 3]), image_token_id=500, video_token_id=501,
        vision_start_token_id=502, vision_end_token_id=503,
        vision_config=dict(depth=2, embed_dim=32, hidden_size=64, hidden_act="quick_gelu", mlp_ratio=2, num_heads=2,
                           in_channels=3, patch_size=14, spatial_merge_size=2, temporal_patch_size=2),
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "granite-tiny": dict(architectures=["GraniteForCausalLM"], model_type="granite", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", rms_norm_eps=1e-6,
        hidden_act="silu", embedding_multiplier=6.0, logits_scaling=4.0, residual_multiplier=0.3, attention_multiplier=0.2,
        rope_theta=10000.0, tie_word_embeddings=True),
    "olmo2-tiny": dict(architectures=["Olmo2ForCausalLM"], model_type="olmo2", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", rms_norm_eps=1e-5, hidden_act="silu",
        rope_theta=10000.0, tie_word_embeddings=False),
    "cohere-tiny": dict(architectures=["CohereForCausalLM"], model_type="cohere", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", layer_norm_eps=1e-5,
        hidden_act="silu", logit_scale=0.25, use_qk_norm=False, rope_theta=10000.0, tie_word_embeddings=True),
    "phi-tiny": dict(architectures=["PhiForCausalLM"], model_type="phi", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", layer_norm_eps=1e-5, hidden_act="gelu_new",
        partial_rotary_factor=0.5, qk_layernorm=False, rope_theta=10000.0, tie_word_embeddings=False),
    "glm4-tiny": dict(architectures=["Glm4ForCausalLM"], model_type="glm4", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", head_dim=16, rms_norm_eps=1e-5,
        hidden_act="silu", attention_bias=True, partial_rotary_factor=0.5, rope_theta=10000.0, tie_word_embeddings=False),
    "nemotron-tiny": dict(architectures=["NemotronForCausalLM"], model_type="nemotron", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", head_dim=16, norm_eps=1e-5,
        hidden_act="relu2", partial_rotary_factor=0.5, rope_theta=10000.0, tie_word_embeddings=False),
    "gemma3-tiny": dict(architectures=["Gemma3ForCausalLM"], model_type="gemma3_text", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", head_dim=16, rms_norm_eps=1e-6,
        hidden_activation="gelu_pytorch_tanh", query_pre_attn_scalar=16, sliding_window=16,
        layer_types=["sliding_attention", "full_attention"],
        rope_parameters=dict(sliding_attention=dict(rope_type="default", rope_theta=10000.0),
                             full_attention=dict(rope_type="default", rope_theta=1000000.0)),
        tie_word_embeddings=True),
    "glm4moe-tiny": dict(architectures=["Glm4MoeForCausalLM"], model_type="glm4_moe", vocab_size=512, hidden_size=64,
        intermediate_size=128, moe_intermediate_size=32, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
        head_dim=16, n_routed_experts=8, n_shared_experts=1, num_experts_per_tok=2, n_group=2, topk_group=1,
        first_k_dense_replace=1, norm_topk_prob=True, routed_scaling_factor=1.5, use_qk_norm=True, attention_bias=True,
        partial_rotary_factor=0.5, rope_theta=10000.0, rms_norm_eps=1e-5, hidden_act="silu", max_position_embeddings=512,
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "llama4-tiny": dict(architectures=["Llama4ForCausalLM"], model_type="llama4_text", vocab_size=512, hidden_size=64,
        intermediate_size=32, intermediate_size_mlp=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2,
        head_dim=16, num_local_experts=4, num_experts_per_tok=1, interleave_moe_layer_step=2, moe_layers=[1, 3],
        no_rope_layers=[1, 1, 1, 0], attention_chunk_size=16, attn_temperature_tuning=True, floor_scale=8, attn_scale=0.1,
        use_qk_norm=True, rms_norm_eps=1e-5, hidden_act="silu", max_position_embeddings=512, attention_bias=False,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0),
        layer_types=["chunked_attention", "chunked_attention", "chunked_attention", "full_attention"],
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "llama4mm-tiny": dict(architectures=["Llama4ForConditionalGeneration"], model_type="llama4", image_token_index=511,
        vision_config=dict(model_type="llama4_vision_model", hidden_size=32, intermediate_size=128, num_hidden_layers=2,
                           num_attention_heads=2, image_size=32, patch_size=8, num_channels=3, vision_output_dim=48,
                           pixel_shuffle_ratio=0.5, projector_input_dim=48, projector_output_dim=48,
                           rope_parameters=dict(rope_type="default", rope_theta=10000.0)),
        vocab_size=512, hidden_size=64, intermediate_size=32, intermediate_size_mlp=128, num_hidden_layers=4, num_attention_heads=4,
        num_key_value_heads=2, head_dim=16, num_local_experts=4, num_experts_per_tok=1, interleave_moe_layer_step=2, moe_layers=[1, 3],
        no_rope_layers=[1, 1, 1, 0], attention_chunk_size=16, attn_temperature_tuning=True, floor_scale=8, attn_scale=0.1,
        use_qk_norm=True, rms_norm_eps=1e-5, hidden_act="silu", max_position_embeddings=512, attention_bias=False,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0),
        layer_types=["chunked_attention", "chunked_attention", "chunked_attention", "full_attention"],
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "qwen25vl-tiny": dict(
        architectures=["Qwen2_5_VLForConditionalGeneration"], model_type="qwen2_5_vl", vocab_size=512, hidden_size=64,
        intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
        max_position_embeddings=512, rms_norm_eps=1e-6, hidden_act="silu", rope_theta=10000.0, attention_bias=True,
        rope_scaling=dict(rope_type="mrope", mrope_section=[2, 3, 3]), image_token_id=500, video_token_id=501,
        vision_start_token_id=502, vision_end_token_id=503,
        vision_config=dict(depth=3, hidden_size=32, out_hidden_size=64, intermediate_size=48, hidden_act="silu", num_heads=2,
                           in_channels=3, patch_size=14, spatial_merge_size=2, temporal_patch_size=2, window_size=56,
                           fullatt_block_indexes=[1], tokens_per_second=2),
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "qwen3vl-tiny": dict(
        architectures=["Qwen3VLForConditionalGeneration"], model_type="qwen3_vl", vocab_size=512, hidden_size=64,
        intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16,
        max_position_embeddings=512, rms_norm_eps=1e-6, hidden_act="silu", rope_theta=10000.0, attention_bias=False,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0, mrope_section=[4, 2, 2], mrope_interleaved=True),
        image_token_id=500, video_token_id=501, vision_start_token_id=502, vision_end_token_id=503,
        vision_config=dict(depth=3, hidden_size=32, out_hidden_size=64, intermediate_size=48, hidden_act="gelu_pytorch_tanh",
                           num_heads=2, in_channels=3, patch_size=16, spatial_merge_size=2, temporal_patch_size=2,
                           num_position_embeddings=36, deepstack_visual_indexes=[0, 1]),
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "qwen3vlmoe-tiny": dict(
        architectures=["Qwen3VLMoeForConditionalGeneration"], model_type="qwen3_vl_moe", vocab_size=512, hidden_size=64,
        intermediate_size=128, moe_intermediate_size=32, num_experts=4, num_experts_per_tok=2, norm_topk_prob=True,
        decoder_sparse_step=1, mlp_only_layers=[], num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2, head_dim=16,
        max_position_embeddings=512, rms_norm_eps=1e-6, hidden_act="silu", rope_theta=10000.0, attention_bias=False,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0, mrope_section=[4, 2, 2], mrope_interleaved=True),
        image_token_id=500, video_token_id=501, vision_start_token_id=502, vision_end_token_id=503,
        vision_config=dict(depth=2, hidden_size=32, out_hidden_size=64, intermediate_size=48, hidden_act="gelu_pytorch_tanh",
                           num_heads=2, in_channels=3, patch_size=16, spatial_merge_size=2, temporal_patch_size=2,
                           num_position_embeddings=36, deepstack_visual_indexes=[0]),
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "qwen35vl-tiny": dict(
        architectures=["Qwen3_5ForConditionalGeneration"], model_type="qwen3_5", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16, linear_num_value_heads=4,
        linear_num_key_heads=2, linear_key_head_dim=16, linear_value_head_dim=16, linear_conv_kernel_dim=4,
        layer_types=["linear_attention", "full_attention", "linear_attention", "full_attention"], rms_norm_eps=1e-6,
        hidden_act="silu", max_position_embeddings=512, attention_bias=False,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0, partial_rotary_factor=0.5, mrope_section=[2, 1, 1],
                             mrope_interleaved=True),
        image_token_id=500, video_token_id=501, vision_start_token_id=502, vision_end_token_id=503,
        vision_config=dict(depth=2, hidden_size=32, out_hidden_size=64, intermediate_size=48, hidden_act="gelu_pytorch_tanh",
                           num_heads=2, in_channels=3, patch_size=16, spatial_merge_size=2, temporal_patch_size=2,
                           num_position_embeddings=36, deepstack_visual_indexes=[]),
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "smollm3-tiny": dict(architectures=["SmolLM3ForCausalLM"], model_type="smollm3", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", rope_theta=10000.0, hidden_act="silu", rms_norm_eps=1e-6,
        no_rope_layers=[1, 1, 1, 0], no_rope_layer_interval=4, tie_word_embeddings=True),
    "cohere2-tiny": dict(architectures=["Cohere2ForCausalLM"], model_type="cohere2", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", rope_theta=10000.0, hidden_act="silu", head_dim=16, layer_norm_eps=1e-5,
        logit_scale=0.25, layer_types=["sliding_attention", "sliding_attention", "sliding_attention", "full_attention"], sliding_window=16, tie_word_embeddings=True),
    "olmo3-tiny": dict(architectures=["Olmo3ForCausalLM"], model_type="olmo3", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", rope_theta=10000.0, hidden_act="silu", rms_norm_eps=1e-5, layer_types=["sliding_attention", "sliding_attention", "sliding_attention", "full_attention"], sliding_window=16,
        tie_word_embeddings=False),
    "exaone4-tiny": dict(architectures=["Exaone4ForCausalLM"], model_type="exaone4", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", rope_theta=10000.0, hidden_act="silu", rms_norm_eps=1e-5, layer_types=["sliding_attention", "sliding_attention", "sliding_attention", "full_attention"], sliding_window=16,
        sliding_window_pattern=4, tie_word_embeddings=False),
    "seedoss-tiny": dict(architectures=["SeedOssForCausalLM"], model_type="seed_oss", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", rope_theta=10000.0, hidden_act="silu", head_dim=16, rms_norm_eps=1e-6,
        attention_bias=True, attention_out_bias=False, tie_word_embeddings=False),
    "arcee-tiny": dict(architectures=["ArceeForCausalLM"], model_type="arcee", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", rope_theta=10000.0, hidden_act="relu2",
        head_dim=16, rms_norm_eps=1e-5, tie_word_embeddings=False),
    "helium-tiny": dict(architectures=["HeliumForCausalLM"], model_type="helium", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", rope_theta=10000.0, hidden_act="silu", head_dim=16, rms_norm_eps=1e-8,
        tie_word_embeddings=False),
    "ernie45-tiny": dict(architectures=["Ernie4_5ForCausalLM"], model_type="ernie4_5", vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", rope_theta=10000.0, hidden_act="silu", head_dim=16, rms_norm_eps=1e-5,
        use_bias=False, tie_word_embeddings=True),
    "olmoe-tiny": dict(architectures=["OlmoeForCausalLM"], model_type="olmoe", vocab_size=512, hidden_size=64,
        intermediate_size=32, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, num_experts=4,
        num_experts_per_tok=2, norm_topk_prob=False, max_position_embeddings=512, rms_norm_eps=1e-5, hidden_act="silu",
        rope_theta=10000.0, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "modernbert-tiny": dict(architectures=["ModernBertModel"], model_type="modernbert", vocab_size=512, hidden_size=64,
        intermediate_size=96, num_hidden_layers=4, num_attention_heads=4, max_position_embeddings=512, norm_eps=1e-5,
        norm_bias=False, attention_bias=False, mlp_bias=False, hidden_activation="gelu", local_attention=8,
        global_attn_every_n_layers=3, layer_types=["full_attention", "sliding_attention", "sliding_attention", "full_attention"],
        rope_parameters=dict(sliding_attention=dict(rope_type="default", rope_theta=10000.0),
                             full_attention=dict(rope_type="default", rope_theta=160000.0)),
        classifier_pooling="cls", pad_token_id=0, bos_token_id=1, eos_token_id=2, cls_token_id=1, sep_token_id=2,
        torch_dtype="float32"),
    "granitemoe-tiny": dict(architectures=["GraniteMoeForCausalLM"], model_type="granitemoe", num_hidden_layers=2, vocab_size=512, hidden_size=64, intermediate_size=32, num_attention_heads=4, num_key_value_heads=2, num_local_experts=4, num_experts_per_tok=2, max_position_embeddings=512, rms_norm_eps=1e-6, hidden_act="silu", rope_theta=10000.0, embedding_multiplier=6.0, logits_scaling=4.0, residual_multiplier=0.3, attention_multiplier=0.2, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=True),
    "granitemoeshared-tiny": dict(architectures=["GraniteMoeSharedForCausalLM"], model_type="granitemoeshared",
        num_hidden_layers=2, shared_intermediate_size=48, vocab_size=512, hidden_size=64, intermediate_size=32, num_attention_heads=4, num_key_value_heads=2, num_local_experts=4, num_experts_per_tok=2, max_position_embeddings=512, rms_norm_eps=1e-6, hidden_act="silu", rope_theta=10000.0, embedding_multiplier=6.0, logits_scaling=4.0, residual_multiplier=0.3, attention_multiplier=0.2, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=True),
    "granitemoehybrid-tiny": dict(architectures=["GraniteMoeHybridForCausalLM"], model_type="granitemoehybrid",
        num_hidden_layers=4, shared_intermediate_size=48, layer_types=["mamba", "attention", "mamba", "attention"],
        position_embedding_type="nope", mamba_n_heads=8, mamba_d_head=16, mamba_expand=2, mamba_d_state=16, mamba_n_groups=2,
        mamba_d_conv=4, mamba_chunk_size=16, mamba_conv_bias=True, mamba_proj_bias=False, vocab_size=512, hidden_size=64, intermediate_size=32, num_attention_heads=4, num_key_value_heads=2, num_local_experts=4, num_experts_per_tok=2, max_position_embeddings=512, rms_norm_eps=1e-6, hidden_act="silu", rope_theta=10000.0, embedding_multiplier=6.0, logits_scaling=4.0, residual_multiplier=0.3, attention_multiplier=0.2, bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=True),
    "gptbigcode-tiny": dict(architectures=["GPTBigCodeForCausalLM"], model_type="gpt_bigcode", vocab_size=512, n_embd=64,
        n_layer=2, n_head=4, n_inner=128, n_positions=512, multi_query=True, activation_function="gelu_pytorch_tanh",
        layer_norm_epsilon=1e-5, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, max_position_embeddings=512,
        bos_token_id=0, eos_token_id=0, torch_dtype="float32", tie_word_embeddings=True),
    "dots1-tiny": dict(architectures=["Dots1ForCausalLM"], model_type="dots1", vocab_size=512, hidden_size=64,
        intermediate_size=128, moe_intermediate_size=32, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
        head_dim=16, n_routed_experts=8, n_shared_experts=1, num_experts_per_tok=2, n_group=2, topk_group=1,
        first_k_dense_replace=1, norm_topk_prob=True, routed_scaling_factor=1.5, attention_bias=False, rope_theta=10000.0,
        rms_norm_eps=1e-5, hidden_act="silu", max_position_embeddings=512, sliding_window=None,
        layer_types=["full_attention", "full_attention", "full_attention"],
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "apertus-tiny": dict(
        architectures=["ApertusForCausalLM"], model_type="apertus", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=256, hidden_act="xielu",
        rms_norm_eps=1e-5, rope_theta=10000.0, rope_parameters=dict(rope_type="default", rope_theta=10000.0), bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32",
        tie_word_embeddings=False),
    "cwm-tiny": dict(
        architectures=["CwmForCausalLM"], model_type="cwm", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=256, sliding_window=16,
        layer_types=["full_attention", "sliding_attention", "sliding_attention", "full_attention"], hidden_act="silu",
        rms_norm_eps=1e-5, rope_theta=10000.0,
        rope_parameters=dict(rope_type="llama3", rope_theta=10000.0, factor=4.0, original_max_position_embeddings=64,
                             low_freq_factor=1.0, high_freq_factor=4.0), bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32",
        tie_word_embeddings=False),
    "jais2-tiny": dict(
        architectures=["Jais2ForCausalLM"], model_type="jais2", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=256,
        hidden_act="relu2", layer_norm_eps=1e-5, attention_bias=True, mlp_bias=True, rope_theta=10000.0, bos_token_id=1,
        eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "ministral3-tiny": dict(
        architectures=["Ministral3ForCausalLM"], model_type="ministral3", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=256,
        sliding_window=None, hidden_act="silu", rms_norm_eps=1e-5, bos_token_id=1, eos_token_id=2, pad_token_id=0,
        rope_parameters=dict(rope_type="yarn", rope_theta=10000.0, factor=4.0, original_max_position_embeddings=16,
                             beta_fast=32.0, beta_slow=1.0, mscale=1.0, mscale_all_dim=1.0, llama_4_scaling_beta=0.1),
        torch_dtype="float32", tie_word_embeddings=False),
    "flexolmo-tiny": dict(
        architectures=["FlexOlmoForCausalLM"], model_type="flex_olmo", vocab_size=512, hidden_size=64, intermediate_size=96,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=256, num_experts=4,
        num_experts_per_tok=2, norm_topk_prob=False, hidden_act="silu", rms_norm_eps=1e-6, rope_theta=10000.0,
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "glm4moelite-tiny": dict(
        architectures=["Glm4MoeLiteForCausalLM"], model_type="glm4_moe_lite", vocab_size=512, hidden_size=64,
        intermediate_size=128, moe_intermediate_size=32, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=4,
        n_shared_experts=1, n_routed_experts=8, routed_scaling_factor=1.8, kv_lora_rank=32, q_lora_rank=48, qk_rope_head_dim=8,
        v_head_dim=16, qk_nope_head_dim=16, n_group=1, topk_group=1, num_experts_per_tok=2, norm_topk_prob=True,
        hidden_act="silu", max_position_embeddings=256, rms_norm_eps=1e-5, rope_interleave=True,
        mlp_layer_types=["dense", "sparse", "sparse"], rope_parameters=dict(rope_type="default", rope_theta=10000.0),
        bos_token_id=0, eos_token_id=1, torch_dtype="float32", tie_word_embeddings=False),
    "dbrx-tiny": dict(
        architectures=["DbrxForCausalLM"], model_type="dbrx", vocab_size=512, d_model=64, n_heads=4, n_layers=2, max_seq_len=256,
        attn_config=dict(kv_n_heads=2, clip_qkv=0.5, rope_theta=10000.0),
        ffn_config=dict(ffn_hidden_size=96, moe_num_experts=4, moe_top_k=2, moe_normalize_expert_weights=1.0,
                        ffn_act_fn=dict(name="silu")),
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "glmmoedsa-tiny": dict(
        architectures=["GlmMoeDsaForCausalLM"], model_type="glm_moe_dsa", vocab_size=512, hidden_size=64,
        intermediate_size=128, moe_intermediate_size=32, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=4,
        n_shared_experts=1, n_routed_experts=8, routed_scaling_factor=1.8, kv_lora_rank=32, q_lora_rank=48, qk_rope_head_dim=8,
        v_head_dim=16, qk_nope_head_dim=16, n_group=1, topk_group=1, num_experts_per_tok=2, norm_topk_prob=True,
        hidden_act="silu", max_position_embeddings=256, rms_norm_eps=1e-5, mlp_layer_types=["dense", "sparse", "sparse"],
        index_topk=256, index_head_dim=16, index_n_heads=2, rope_parameters=dict(rope_type="default", rope_theta=10000.0),
        bos_token_id=0, eos_token_id=1, torch_dtype="float32", tie_word_embeddings=False),
    "afmoe-tiny": dict(
        architectures=["AfmoeForCausalLM"], model_type="afmoe", vocab_size=512, hidden_size=64, intermediate_size=128,
        moe_intermediate_size=32, num_hidden_layers=4, num_dense_layers=1, num_attention_heads=4, num_key_value_heads=2,
        head_dim=16, max_position_embeddings=256, num_experts=8, num_experts_per_tok=2, num_shared_experts=1, route_scale=1.5,
        global_attn_every_n_layers=2, sliding_window=16, mup_enabled=True, hidden_act="silu", rms_norm_eps=1e-5,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0), bos_token_id=1, eos_token_id=2, pad_token_id=0,
        torch_dtype="float32", tie_word_embeddings=False),
    "minimaxm2-tiny": dict(
        architectures=["MiniMaxM2ForCausalLM"], model_type="minimax_m2", vocab_size=512, hidden_size=64, intermediate_size=96,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=256,
        num_local_experts=4, num_experts_per_tok=2, hidden_act="silu", rms_norm_eps=1e-6,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0, partial_rotary_factor=0.5), bos_token_id=1, eos_token_id=2,
        pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "exaonemoe-tiny": dict(
        architectures=["ExaoneMoeForCausalLM"], model_type="exaone_moe", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=256,
        sliding_window=16, sliding_window_pattern=2, first_k_dense_replace=1, moe_intermediate_size=32, num_experts=8,
        num_experts_per_tok=2, num_shared_experts=1, norm_topk_prob=True, routed_scaling_factor=2.5, n_group=2, topk_group=1,
        hidden_act="silu", rms_norm_eps=1e-5, rope_parameters=dict(rope_type="default", rope_theta=10000.0), bos_token_id=1,
        eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "solaropen-tiny": dict(
        architectures=["SolarOpenForCausalLM"], model_type="solar_open", vocab_size=512, hidden_size=64, num_hidden_layers=2,
        num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=256, moe_intermediate_size=32,
        num_experts_per_tok=2, n_shared_experts=1, n_routed_experts=8, routed_scaling_factor=1.5, n_group=2, topk_group=1,
        norm_topk_prob=True, hidden_act="silu", rms_norm_eps=1e-5, attention_bias=False,
        rope_parameters=dict(rope_type="default", rope_theta=10000.0, partial_rotary_factor=1.0), bos_token_id=1, eos_token_id=2,
        pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "hunyuan-dense-tiny": dict(   # q/k norms after the rotation
        architectures=["HunYuanDenseV1ForCausalLM"], model_type="hunyuan_v1_dense", vocab_size=512, hidden_size=64,
        intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=16,
        max_position_embeddings=256, hidden_act="silu", rms_norm_eps=1e-5, rope_theta=10000.0, use_qk_norm=True,
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "phimoe-tiny": dict(
        architectures=["PhimoeForCausalLM"], model_type="phimoe", vocab_size=512, hidden_size=64, intermediate_size=96,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, num_local_experts=4, num_experts_per_tok=2,
        max_position_embeddings=256, rms_norm_eps=1e-5, rope_theta=10000.0, hidden_act="silu", attention_bias=True,
        lm_head_bias=True, router_jitter_noise=0.01, input_jitter_noise=0.0, sliding_window=None,
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "ministral-tiny": dict(
        architectures=["MinistralForCausalLM"], model_type="ministral", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=256,
        sliding_window=16, layer_types=["sliding_attention", "full_attention", "sliding_attention", "full_attention"],
        rms_norm_eps=1e-5, rope_theta=10000.0, hidden_act="silu", bos_token_id=1, eos_token_id=2, torch_dtype="float32",
        tie_word_embeddings=False),
    "glm-tiny": dict(
        architectures=["GlmForCausalLM"], model_type="glm", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=256,
        partial_rotary_factor=0.5, attention_bias=True, rms_norm_eps=1e-5, rope_theta=10000.0, hidden_act="silu",
        bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float32", tie_word_embeddings=False),
    "olmo-tiny": dict(
        architectures=["OlmoForCausalLM"], model_type="olmo", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=256,
        rope_theta=10000.0, hidden_act="silu", clip_qkv=2.0, bos_token_id=1, eos_token_id=2, pad_token_id=0,
        torch_dtype="float32", tie_word_embeddings=False),
    "stablelm-tiny": dict(
        architectures=["StableLmForCausalLM"], model_type="stablelm", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=256,
        partial_rotary_factor=0.25, rope_theta=10000.0, hidden_act="silu", layer_norm_eps=1e-5, use_qkv_bias=True,
        bos_token_id=1, eos_token_id=2, torch_dtype="float32", tie_word_embeddings=False),
    "jamba-tiny": dict(
        architectures=["JambaForCausalLM"], model_type="jamba", vocab_size=512, hidden_size=64, intermediate_size=128,
        num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, mamba_d_state=8, mamba_d_conv=4,
        mamba_expand=2, mamba_dt_rank=4, attn_layer_period=2, attn_layer_offset=1, expert_layer_period=2,
        expert_layer_offset=1, num_experts=4, num_experts_per_tok=2, use_mamba_kernels=False, rms_norm_eps=1e-6,
        max_position_embeddings=256, hidden_act="silu", bos_token_id=1, eos_token_id=2, pad_token_id=0,
        torch_dtype="float32", tie_word_embeddings=False),
    "bart-tiny": dict(
        architectures=["BartForConditionalGeneration"], model_type="bart", vocab_size=512, d_model=64, hidden_size=64,
        encoder_layers=2, decoder_layers=2, num_hidden_layers=2, encoder_attention_heads=4, decoder_attention_heads=4,
        num_attention_heads=4, encoder_ffn_dim=128, decoder_ffn_dim=128, max_position_embeddings=128,
        activation_function="gelu", scale_embedding=False, pad_token_id=1, bos_token_id=0, eos_token_id=2,
        decoder_start_token_id=2, torch_dtype="float32", tie_word_embeddings=True, is_encoder_decoder=True),
    "whisper-tiny-test": dict(
        architectures=["WhisperForConditionalGeneration"], model_type="whisper", vocab_size=512, d_model=64,
        hidden_size=64, encoder_layers=2, decoder_layers=2, num_hidden_layers=2, encoder_attention_heads=4,
        decoder_attention_heads=4, num_attention_heads=4, encoder_ffn_dim=128, decoder_ffn_dim=128, num_mel_bins=80,
        max_source_positions=1500, max_target_positions=128, max_position_embeddings=128,
        activation_function="gelu", scale_embedding=False, pad_token_id=0, bos_token_id=1, eos_token_id=2,
        decoder_start_token_id=1, torch_dtype="float32", tie_word_embeddings=True, is_encoder_decoder=True),
    "nomicbert-tiny": dict(
        architectures=["NomicBertModel"], model_type="nomic_bert", vocab_size=512, hidden_size=64, num_hidden_layers=2,
        num_attention_heads=4, head_dim=16, intermediate_size=128, max_position_embeddings=256, type_vocab_size=2, hidden_act="silu",
        layer_norm_eps=1e-12, rope_parameters=dict(rope_type="default", rope_theta=1000.0), torch_dtype="float32", pad_token_id=0),
    "bert-tiny": dict(
        architectures=["BertModel"], model_type="bert", vocab_size=512, hidden_size=64, num_hidden_layers=2,
        num_attention_heads=4, intermediate_size=128, max_position_embeddings=256, type_vocab_size=2,
        layer_norm_eps=1e-12, hidden_act="gelu", torch_dtype="float32", pad_token_id=0),
    "qwen2moe-tiny": dict(
        architectures=["Qwen2MoeForCausalLM"], model_type="qwen2_moe", vocab_size=512, hidden_size=128,
        intermediate_size=256, moe_intermediate_size=64, shared_expert_intermediate_size=128,
        num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=1024,
        rms_norm_eps=1e-6, rope_theta=10000.0, hidden_act="silu", num_experts=4, num_experts_per_tok=2,
        norm_topk_prob=False, decoder_sparse_step=1, mlp_only_layers=[], bos_token_id=1, eos_token_id=2,
        torch_dtype="float32", tie_word_embeddings=False),
    "qwen3moe-tiny": dict(
        architectures=["Qwen3MoeForCausalLM"], model_type="qwen3_moe", vocab_size=512, hidden_size=128,
        intermediate_size=256, moe_intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=2, head_dim=32, max_position_embeddings=1024, rms_norm_eps=1e-6,
        rope_theta=10000.0, hidden_act="silu", num_experts=4, num_experts_per_tok=2, norm_topk_prob=True,
        decoder_sparse_step=1, mlp_only_layers=[], bos_token_id=1, eos_token_id=2, torch_dtype="float32",
        tie_word_embeddings=False),
    "deepseek-v3": dict(
        architectures=["DeepseekV3ForCausalLM"], model_type="deepseek_v3", vocab_size=129280,
        hidden_size=7168, intermediate_size=18432, moe_intermediate_size=2048, num_hidden_layers=61,
        num_attention_heads=128, num_key_value_heads=128, q_lora_rank=1536, kv_lora_rank=512,
        qk_nope_head_dim=128, qk_rope_head_dim=64, v_head_dim=128, n_routed_experts=256,
        n_shared_experts=1, num_experts_per_tok=8, n_group=8, topk_group=4, routed_scaling_factor=2.5,
        first_k_dense_replace=3, moe_layer_freq=1, norm_topk_prob=True, scoring_func="sigmoid",
        topk_method="noaux_tc", max_position_embeddings=163840, rms_norm_eps=1e-6, rope_theta=10000.0,
        rope_scaling=dict(type="yarn", factor=40, original_max_position_embeddings=4096,
                          beta_fast=32, beta_slow=1, mscale=1.0, mscale_all_dim=1.0),
        hidden_act="silu", bos_token_id=0, eos_token_id=1, torch_dtype="bfloat16",
        tie_word_embeddings=False,
        quantization_config=dict(quant_method="fp8", fmt="e4m3", weight_block_size=[128, 128],
                                 activation_scheme="dynamic")),
    "deepseek-tiny": dict(
        architectures=["DeepseekV3ForCausalLM"], model_type="deepseek_v3", vocab_size=512,
        hidden_size=128, intermediate_size=256, moe_intermediate_size=64, num_hidden_layers=3,
        num_attention_heads=4, num_key_value_heads=4, q_lora_rank=48, kv_lora_rank=32,
        qk_nope_head_dim=32, qk_rope_head_dim=16, v_head_dim=32, n_routed_experts=8,
        n_shared_experts=1, num_experts_per_tok=2, n_group=2, topk_group=1, routed_scaling_factor=2.5,
        first_k_dense_replace=1, moe_layer_freq=1, norm_topk_prob=True, scoring_func="sigmoid",
        topk_method="noaux_tc", max_position_embeddings=1024, rms_norm_eps=1e-6, rope_theta=10000.0,
        hidden_act="silu", bos_token_id=0, eos_token_id=1, torch_dtype="float32",
        tie_word_embeddings=False),
    "deepseek-small": dict(
        architectures=["DeepseekV3ForCausalLM"], model_type="deepseek_v3", vocab_size=4096,
        hidden_size=1024, intermediate_size=2048, moe_intermediate_size=512, num_hidden_layers=3,
        num_attention_heads=16, num_key_value_heads=16, q_lora_rank=384, kv_lora_rank=512,
        qk_nope_head_dim=128, qk_rope_head_dim=64, v_head_dim=128, n_routed_experts=16,
        n_shared_experts=1, num_experts_per_tok=4, n_group=4, topk_group=2, routed_scaling_factor=2.5,
        first_k_dense_replace=1, moe_layer_freq=1, norm_topk_prob=True, scoring_func="sigmoid",
        topk_method="noaux_tc", max_position_embeddings=4096, rms_norm_eps=1e-6, rope_theta=10000.0,
        hidden_act="silu", bos_token_id=0, eos_token_id=1, torch_dtype="bfloat16",
        tie_word_embeddings=False),
}

_ALIASES = {
    "facebook/opt-125m": "opt-125m",
    "meta-llama/meta-llama-3-8b": "llama-3-8b", "meta-llama/meta-llama-3-8b-instruct": "llama-3-8b",
    "meta-llama/llama-3.1-8b": "llama-3-8b", "meta-llama/llama-3.1-8b-instruct": "llama-3-8b",
    "meta-llama/meta-llama-3-70b": "llama-3-70b", "meta-llama/meta-llama-3-70b-instruct": "llama-3-70b",
    "meta-llama/llama-3.1-70b": "llama-3-70b", "meta-llama/llama-3.1-70b-instruct": "llama-3-70b",
    "mistralai/mixtral-8x7b-v0.1": "mixtral-8x7b", "mistralai/mixtral-8x7b-instruct-v0.1": "mixtral-8x7b",
    "deepseek-ai/deepseek-v3": "deepseek-v3", "openai-community/gpt2": "gpt2",
    "mistralai/mistral-7b-v0.1": "mistral-7b", "qwen/qwen2.5-7b": "qwen2.5-7b", "qwen/qwen3-8b": "qwen3-8b",
    "google/gemma-2b": "gemma-2b",
}


def get_preset_hf_config(name: str) -> Optional[dict[str, Any]]:
    key = name.lower()
    key = _ALIASES.get(key, key)
    cfg = PRESETS.get(key)
    return dict(cfg) if cfg is not None else None


def _flatten_nested_config(c: dict) -> None:
    """Families whose config.json nests or renames the fields the engine sizes the KV cache from."""
    if c.get("model_type") == "dbrx":
        a, f = c.get("attn_config") or {}, c.get("ffn_config") or {}
        c.setdefault("hidden_size", c.get("d_model"))
        c.setdefault("num_attention_heads", c.get("n_heads"))
        c.setdefault("num_hidden_layers", c.get("n_layers"))
        c.setdefault("max_position_embeddings", c.get("max_seq_len", 2048))
        c.setdefault("num_key_value_heads", a.get("kv_n_heads") or c.get("n_heads"))
        c.setdefault("rope_theta", a.get("rope_theta", 10000.0))
        c.setdefault("clip_qkv", a.get("clip_qkv"))
        c.setdefault("num_local_experts", f.get("moe_num_experts", 1))
        c.setdefault("num_experts_per_tok", f.get("moe_top_k", 1))
        c.setdefault("hidden_act", (f.get("ffn_act_fn") or {}).get("name", "silu"))
        if "intermediate_size" not in c:
            # released checkpoints: ffn_hidden_size is the expert width; transformers >= 5 swaps the two field names
            fh, hs = f.get("ffn_hidden_size"), f.get("hidden_size")
            c["intermediate_size"] = hs if (hs and fh == c.get("d_model") and hs != fh) else fh


@dataclass
class ModelConfig:
    model: str = "opt-tiny"
    tokenizer: Optional[str] = None
    tokenizer_mode: str = "auto"
    trust_remote_code: bool = False
    dtype: str = "auto"
    seed: int = 0
    revision: Optional[str] = None
    max_model_len: int = 0  # 0 → derive from the architecture
    quantization: Optional[str] = None  # None|fp8|fp8_block|nvfp4|awq|gptq|...
    enforce_eager: bool = False
    max_logprobs: int = 20
    served_model_name: Optional[str] = None
    hf_overrides: dict[str, Any] = field(default_factory=dict)
    runner: str = "generate"  # generate|pooling
    model_impl: str = "auto"  # auto|vllm_b200|transformers (reference: ``--model-impl``)
    skip_tokenizer_init: bool = False
    generation_config: str = "auto"
    logits_processors: Optional[list[str]] = None
    # resolved
    hf_config: dict[str, Any] = field(default_factory=dict)

    def __post_init__(self) -> None:
        if not self.hf_config:
            self.hf_config = self._load_hf_config()
        if self.hf_config.get("model_type") == "ultravox" and "projector_hidden_size" not in self.hf_config:
            # Ultravox's top-level ``hidden_size`` is the projector width, not the decoder's
            self.hf_config["projector_hidden_size"] = self.hf_config.pop("hidden_size", 4096)
        if isinstance(self.hf_config.get("text_config"), dict):
            # vision-language checkpoints nest the decoder's fields: lift them to the top level (top-level wins)
            tc = self.hf_config.pop("text_config")
            self.hf_config = {**tc, **self.hf_config}
            if tc.get("model_type"):
                self.hf_config.setdefault("text_model_type", tc["model_type"])
        self.hf_config.update(self.hf_overrides)
        _flatten_nested_config(self.hf_config)
        if self.tokenizer is None:
            self.tokenizer = self.model
        if self.served_model_name is None:
            self.served_model_name = self.model
        derived = self._derive_max_len()
        if self.max_model_len <= 0:
            self.max_model_len = derived
        if self.quantization is None:
            qc = self.hf_config.get("quantization_config")
            if qc:
                m = qc.get("quant_method")
                if m == "fp8" and qc.get("weight_block_size"):
                    self.quantization = "fp8_block"
                else:
                    self.quantization = m

    # -- loading ----------------------------------------------------------
    def _load_hf_config(self) -> dict[str, Any]:
        if self.model.endswith(".gguf") and os.path.isfile(self.model):
            from vllm_b200.model_loader.gguf import gguf_config
            return gguf_config(self.model)
        path = os.path.join(self.model, "config.json") if os.path.isdir(self.model) else None
        if path and os.path.exists(path):
            with open(path) as f:
                return json.load(f)
        cfg = get_preset_hf_config(self.model)
        if cfg is None:
            raise ValueError(
                f"model {self.model!r} is neither a local directory with config.json nor a built-in "
                f"preset ({sorted(PRESETS)})")
        return cfg

    def _derive_max_len(self) -> int:
        c = self.hf_config
        for k in ("max_position_embeddings", "n_positions", "max_sequence_length", "seq_length"):
            if c.get(k):
                return int(c[k])
        return 2048

    # -- accessors --------------------------------------------------------
    @property
    def is_encoder_only(self) -> bool:
        return (self.architecture.startswith(("Bert", "Roberta", "XLMRoberta", "NomicBert", "HF_ColBERT", "ColBERTJinaRoberta", "BgeM3"))
                or "Bidirectional" in self.architecture
                or (self.hf_config.get("is_causal") is False and not self.prefix_lm))

    @property
    def prefix_lm(self) -> bool:
        """Generative models whose PROMPT attends bidirectionally (PaliGemma): the prompt must be one prefill pass."""
        return self.architecture.startswith("PaliGemma")

    @property
    def architecture(self) -> str:
        return self.hf_config.get("architectures", ["LlamaForCausalLM"])[0]

    @property
    def torch_dtype(self) -> torch.dtype:
        d = _DTYPES.get(self.dtype, None) if self.dtype in _DTYPES else None
        if self.dtype != "auto" and d is None:
            raise ValueError(f"unknown dtype {self.dtype}")
        if d is not None:
            return d
        hd = _DTYPES.get(str(self.hf_config.get("torch_dtype", "bfloat16")), torch.bfloat16)
        if hd == torch.float16:
            return torch.float16
        return hd or torch.bfloat16

    @property
    def vocab_size(self) -> int:
        return int(self.hf_config["vocab_size"])

    @property
    def hidden_size(self) -> int:
        c = self.hf_config
        return int(c.get("hidden_size") or c.get("n_embd") or c.get("d_model") or 0)

    @property
    def num_layers(self) -> int:
        c = self.hf_config
        return int(c.get("num_hidden_layers") or c.get("n_layer") or c.get("num_layers") or c.get("n_layers") or 0)

    @property
    def num_attention_heads(self) -> int:
        c = self.hf_config
        return int(c.get("num_attention_heads") or c.get("n_head") or c.get("num_heads") or c.get("n_heads") or 0)

    @property
    def num_kv_heads(self) -> int:
        if self.is_mla:
            return 1
        return int(self.hf_config.get("num_key_value_heads") or self.num_attention_heads)

    @property
    def head_dim(self) -> int:
        if self.is_mla:
            return int(self.hf_config["kv_lora_rank"]) + int(self.hf_config["qk_rope_head_dim"])
        hd = self.hf_config.get("head_dim")
        return int(hd) if hd else self.hidden_size // self.num_attention_heads

    @property
    def is_mla(self) -> bool:
        return self.hf_config.get("kv_lora_rank") is not None

    @property
    def is_moe(self) -> bool:
        return any(self.hf_config.get(k) for k in ("num_local_experts", "n_routed_experts", "num_experts"))

    @property
    def sliding_window(self) -> Optional[int]:
        return self.hf_config.get("sliding_window")

    @property
    def eos_token_id(self) -> Optional[int]:
        e = self.hf_config.get("eos_token_id")
        if isinstance(e, list):
            return e[0]
        return e

    def get_num_kv_heads(self, tp: int) -> int:
        return max(1, self.num_kv_heads // tp)

    def get_num_attention_heads(self, tp: int) -> int:
        return self.num_attention_heads // tp

    def get_num_layers(self, pp_rank: int = 0, pp_size: int = 1) -> int:
        from vllm_b200.parallel.utils import get_pp_indices
        s, e = get_pp_indices(self.num_layers, pp_rank, pp_size)
        return e - s

    def kv_bytes_per_token_per_layer(self, tp: int, kv_dtype_size: int) -> int:
        if self.is_mla:
            return self.head_dim * kv_dtype_size
        return 2 * self.get_num_kv_heads(tp) * self.head_dim * kv_dtype_size

    def verify_with_parallel_config(self, pc) -> None:
        tp = pc.tensor_parallel_size
        if self.num_attention_heads % tp != 0:
            raise ValueError(
                f"num_attention_heads ({self.num_attention_heads}) must be divisible by TP size ({tp})")
        if pc.pipeline_parallel_size > self.num_layers:
            raise ValueError("pipeline_parallel_size exceeds the number of layers")

    def compute_hash(self) -> str:
        import hashlib
        return hashlib.sha256(json.dumps(
            [self.model, self.dtype, self.quantization, self.hf_config], sort_keys=True, default=str
        ).encode()).hexdigest()[:16]
