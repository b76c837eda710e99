"""End-to-end greedy equivalence with HuggingFace on tiny random-init models (CPU reference path).
Mirrors the reference's tests/models/** strategy (HfRunner vs VllmRunner, greedy match)."""
import pytest
import torch

from tests.utils import hf_greedy, make_hf_model, make_llm
from vllm_b200 import SamplingParams

PROMPTS = [[5, 17, 99, 3, 250, 41], [7, 7, 7, 301, 12], list(range(10, 60))]


@pytest.mark.parametrize("preset", ["llama-tiny", "opt-tiny", "gpt2-tiny", "mixtral-tiny", "qwen2moe-tiny",
                                    "qwen3moe-tiny", "deepseek-tiny", "mamba-tiny", "phi3-tiny", "gemma2-tiny", "gptneox-tiny",
                                    "starcoder2-tiny", "gptj-tiny", "olmo-tiny", "stablelm-tiny", "jamba-tiny",
                                    "bloom-tiny", "mpt-tiny", "falcon-rw-tiny", "falcon-7b-tiny", "falcon-40b-tiny",
                                    "gptoss-tiny", "mamba2-tiny", "bamba-tiny", "qwen3next-tiny",
                                    "granite-tiny", "olmo2-tiny", "cohere-tiny", "phi-tiny", "glm4-tiny", "nemotron-tiny",
                                    "gemma3-tiny", "glm4moe-tiny", "llama4-tiny",
                                    "smollm3-tiny", "cohere2-tiny", "olmo3-tiny", "exaone4-tiny", "seedoss-tiny", "arcee-tiny",
                                    "helium-tiny", "ernie45-tiny", "olmoe-tiny", "granitemoe-tiny", "granitemoeshared-tiny",
                                    "granitemoehybrid-tiny", "gptbigcode-tiny", "dots1-tiny", "ministral-tiny", "glm-tiny",
                                    "apertus-tiny", "hunyuan-dense-tiny", "phimoe-tiny", "falconmamba-tiny", "biogpt-tiny", "xglm-tiny", "vaultgemma-tiny", "ernie45moe-tiny", "hunyuan-moe-tiny",
                                    "cwm-tiny", "jais2-tiny", "ministral3-tiny", "flexolmo-tiny",
                                    "minimaxm2-tiny", "exaonemoe-tiny", "solaropen-tiny", "afmoe-tiny", "glm4moelite-tiny",
                                    "qwen35-tiny", "qwen35moe-tiny", "nemotronh-tiny", "nemotronh-latent-tiny",
                                    "falconh1-tiny", "falconh1-nonorm-tiny", "lfm2-tiny", "lfm2moe-tiny", "olmohybrid-tiny"])
def test_greedy_matches_hf(preset):
    hf = make_hf_model(preset)
    llm = make_llm(preset, weights=hf.state_dict().items())
    sp = SamplingParams(max_tokens=12, temperature=0.0, ignore_eos=True)
    outs = llm.generate([{"prompt_token_ids": p} for p in PROMPTS], sp)
    for p, o in zip(PROMPTS, outs):
        got, exp = o.outputs[0].token_ids, hf_greedy(hf, p, 12)
        if got != exp:
            # a (near-)tie in the reference logits is not a disagreement: accept a mismatch only there (fp32 noise ~1e-5)
            k = next(i for i, (a, b) in enumerate(zip(got, exp)) if a != b)
            with torch.no_grad():
                lg = hf(torch.tensor([p + exp[:k]])).logits[0, -1]
            assert abs(float(lg[got[k]] - lg[exp[k]])) < 2e-4, (preset, k, got, exp)
    llm.shutdown()


def test_chunked_prefill_and_prefix_cache_equivalence():
    hf = make_hf_model("llama-tiny")
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    long_prompt = [(i * 7 + 3) % 500 + 3 for i in range(150)]
    ref = hf_greedy(hf, long_prompt, 8)
    # chunked: budget 32 tokens per step forces 5 prefill chunks
    llm = make_llm("llama-tiny", weights=hf.state_dict().items(), max_num_batched_tokens=32)
    o1 = llm.generate([{"prompt_token_ids": long_prompt}], sp)[0]
    assert o1.outputs[0].token_ids == ref
    # second time hits the prefix cache (all full blocks but the last token's)
    o2 = llm.generate([{"prompt_token_ids": long_prompt}], sp)[0]
    assert o2.outputs[0].token_ids == ref
    assert o2.num_cached_tokens >= 128
    llm.shutdown()


def test_logprobs_and_prompt_logprobs():
    hf = make_hf_model("llama-tiny")
    llm = make_llm("llama-tiny", weights=hf.state_dict().items())
    p = PROMPTS[0]
    sp = SamplingParams(max_tokens=4, temperature=0.0, logprobs=3, prompt_logprobs=2, ignore_eos=True)
    o = llm.generate([{"prompt_token_ids": p}], sp)[0]
    with torch.no_grad():
        full = torch.tensor([p + o.outputs[0].token_ids])
        lp = hf(full).logits[0].log_softmax(-1)
    for i, (tok, d) in enumerate(zip(o.outputs[0].token_ids, o.outputs[0].logprobs)):
        assert abs(d[tok].logprob - lp[len(p) - 1 + i, tok].item()) < 1e-3
        assert d[tok].rank == 1
        assert len(d) in (3, 4)
    assert o.prompt_logprobs[0] is None and len(o.prompt_logprobs) == len(p)
    for i in range(1, len(p)):
        assert abs(o.prompt_logprobs[i][p[i]].logprob - lp[i - 1, p[i]].item()) < 1e-3
    llm.shutdown()


def test_sampling_params_behaviour():
    llm = make_llm("llama-tiny")
    # n>1 with a seed is reproducible; different children differ
    sp = SamplingParams(n=3, max_tokens=6, temperature=1.0, seed=7, ignore_eos=True)
    a = llm.generate(["abc"], sp)[0]
    b = llm.generate(["abc"], sp)[0]
    assert len(a.outputs) == 3
    assert [c.token_ids for c in a.outputs] == [c.token_ids for c in b.outputs]
    # stop token ids / max_tokens / min_tokens
    g = llm.generate(["abc"], SamplingParams(max_tokens=10, temperature=0, ignore_eos=True))[0].outputs[0].token_ids
    st = llm.generate(["abc"], SamplingParams(max_tokens=10, temperature=0, stop_token_ids=[g[3]]))[0].outputs[0]
    assert st.token_ids == g[:4] and st.finish_reason == "stop" and st.stop_reason == g[3]
    # allowed_token_ids and logit_bias steer the argmax
    al = llm.generate(["abc"], SamplingParams(max_tokens=4, temperature=0, allowed_token_ids=[11, 12]))[0]
    assert set(al.outputs[0].token_ids) <= {11, 12}
    lb = llm.generate(["abc"], SamplingParams(max_tokens=3, temperature=0, logit_bias={77: 100.0}, ignore_eos=True))[0]
    assert lb.outputs[0].token_ids == [77, 77, 77]
    # top_k=1 equals greedy even at high temperature
    tk = llm.generate(["abc"], SamplingParams(max_tokens=10, temperature=5.0, top_k=1, ignore_eos=True))[0]
    assert tk.outputs[0].token_ids == g
    # repetition penalty changes the continuation of a repetitive greedy output
    llm.shutdown()


def test_stop_strings_and_detokenizer():
    llm = make_llm("llama-tiny")
    base = llm.generate(["hello"], SamplingParams(max_tokens=24, temperature=0, ignore_eos=True))[0].outputs[0]
    text = base.text
    if len(text) > 6:
        stop = text[3:5]
        o = llm.generate(["hello"], SamplingParams(max_tokens=24, temperature=0, stop=[stop], ignore_eos=True))[0].outputs[0]
        assert o.finish_reason == "stop" and o.stop_reason == stop
        assert o.text == text[:text.find(stop)]
        o2 = llm.generate(["hello"], SamplingParams(max_tokens=24, temperature=0, stop=[stop], ignore_eos=True,
                                                     include_stop_str_in_output=True))[0].outputs[0]
        assert o2.text == text[:text.find(stop) + len(stop)]
    llm.shutdown()


def test_async_scheduling_matches_sync():
    """Async scheduling (step N+1 launched before step N's tokens reach the host) must not change outputs."""
    hf = make_hf_model("llama-tiny")
    sp = SamplingParams(max_tokens=10, temperature=0.0, ignore_eos=True)
    prompts = [{"prompt_token_ids": p} for p in PROMPTS]
    a = make_llm("llama-tiny", weights=hf.state_dict().items())
    ref = [o.outputs[0].token_ids for o in a.generate(prompts, sp)]
    a.shutdown()
    b = make_llm("llama-tiny", weights=hf.state_dict().items(), async_scheduling=True)
    got = [o.outputs[0].token_ids for o in b.generate(prompts, sp)]
    # stop conditions discovered one step late must still cut at the right place
    eos_sp = SamplingParams(max_tokens=10, temperature=0.0, stop_token_ids=[ref[0][4]])
    cut = b.generate(prompts[:1], eos_sp)[0].outputs[0].token_ids
    b.shutdown()
    assert got == ref
    assert cut == ref[0][:ref[0].index(ref[0][4]) + 1]


def test_bert_embeddings_match_hf():
    """Encoder-only model through the pooling runner: [CLS] embedding == HF BertModel, bidirectional attention."""
    hf = make_hf_model("bert-tiny")
    llm = make_llm("bert-tiny", weights=hf.state_dict().items())
    prompts = [[3, 17, 99, 250, 41, 7], list(range(10, 45))]
    outs = llm.embed([{"prompt_token_ids": p} for p in prompts])
    for p, o in zip(prompts, outs):
        with torch.no_grad():
            ref = hf(torch.tensor([p])).last_hidden_state[0, 0]
        ref = torch.nn.functional.normalize(ref, dim=-1)
        got = torch.as_tensor(o.outputs.data).float()
        torch.testing.assert_close(got, ref, atol=2e-4, rtol=1e-3)
    llm.shutdown()


def test_llava_matches_hf_with_image():
    """Vision tower + projector + placeholder merge + decoder == HF LlavaForConditionalGeneration (greedy)."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llava-tiny")
    text = {k: v for k, v in cfg.items() if k not in ("architectures", "model_type", "vision_config", "image_token_index",
                                                        "vision_feature_layer", "vision_feature_select_strategy",
                                                        "projector_hidden_act", "torch_dtype")}
    hf_cfg = transformers.LlavaConfig(vision_config=cfg["vision_config"], text_config=dict(model_type="llama", **text),
                                      image_token_index=511, vision_feature_layer=-2,
                                      vision_feature_select_strategy="default", projector_hidden_act="gelu")
    torch.manual_seed(0)
    hf = transformers.LlavaForConditionalGeneration(hf_cfg).eval()
    llm = make_llm("llava-tiny", weights=hf.state_dict().items())
    g = torch.Generator().manual_seed(1)
    img = torch.randn(3, 28, 28, generator=g)
    ids = [5, 17, 511, 99, 3, 250]                       # one <image> placeholder → 4 patch tokens
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    out = llm.generate([{"prompt_token_ids": ids, "multi_modal_data": {"image": img}},
                        {"prompt_token_ids": [7, 8, 9, 10]}], sp)
    got = out[0].outputs[0].token_ids
    exp_ids = [5, 17] + [511] * 4 + [99, 3, 250]
    with torch.no_grad():
        ref = hf.generate(input_ids=torch.tensor([exp_ids]), pixel_values=img[None], max_new_tokens=8, do_sample=False,
                          pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
    assert got == ref
    # the same image again: served from the encoder cache
    runner = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner
    llm.generate([{"prompt_token_ids": [1, 511, 2], "multi_modal_data": {"image": img}}], sp)
    assert runner.encoder_cache.hits >= 1
    llm.shutdown()


def test_bart_encoder_decoder_matches_hf():
    """Seq2seq: encoder once per request, cross-attention K/V kept per request, causal decoder over the paged cache."""
    hf = make_hf_model("bart-tiny")
    llm = make_llm("bart-tiny", weights=hf.state_dict().items(), enable_prefix_caching=False)
    srcs = [[0, 17, 99, 3, 250, 41, 2], [0] + list(range(10, 40)) + [2]]
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    outs = llm.generate([{"encoder_prompt_token_ids": s} for s in srcs], sp)
    for s, o in zip(srcs, outs):
        with torch.no_grad():
            ref = hf.generate(input_ids=torch.tensor([s]), max_new_tokens=8, do_sample=False, num_beams=1,
                              decoder_start_token_id=2, eos_token_id=None, pad_token_id=1,
                              forced_bos_token_id=None, forced_eos_token_id=None)[0, 1:].tolist()
        assert o.outputs[0].token_ids == ref[:8], (o.outputs[0].token_ids, ref)
    llm.shutdown()


def test_whisper_matches_hf_and_feature_extractor():
    """Speech path: our log-mel front-end == HF WhisperFeatureExtractor, and waveform → encoder → cross-attending
    decoder reproduces the HF model's greedy continuation."""
    import numpy as np
    from transformers import WhisperFeatureExtractor
    from vllm_b200.multimodal.audio import log_mel_spectrogram
    rng = np.random.default_rng(0)
    t = np.arange(16000 * 3) / 16000
    waves = [(0.3 * np.sin(2 * np.pi * 440 * t) + 0.05 * rng.standard_normal(t.shape)).astype(np.float32),
             (0.2 * np.sin(2 * np.pi * 220 * t[:20000]) * np.hanning(20000)).astype(np.float32)]
    fe = WhisperFeatureExtractor(feature_size=80)
    feats = [fe(w, sampling_rate=16000, return_tensors="pt").input_features[0] for w in waves]
    for w, f in zip(waves, feats):
        assert (log_mel_spectrogram(w, 80) - f).abs().max() < 1e-4

    hf = make_hf_model("whisper-tiny-test")
    llm = make_llm("whisper-tiny-test", weights=hf.state_dict().items(), enable_prefix_caching=False, max_model_len=128)
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    prompts = [{"prompt_token_ids": [1, 7, 9], "multi_modal_data": {"audio": (waves[0], 16000)}},
               {"multi_modal_data": {"audio": (waves[1], 16000)}}]
    outs = llm.generate(prompts, sp)
    for p, f, o in zip(prompts, feats, outs):
        dec = list(p.get("prompt_token_ids", [1]))
        with torch.no_grad():
            enc = hf.model.encoder(f[None]).last_hidden_state
            for _ in range(8):
                h = hf.model.decoder(input_ids=torch.tensor([dec]), encoder_hidden_states=enc).last_hidden_state
                dec.append(int(hf.proj_out(h[0, -1]).argmax()))
        n0 = len(p.get("prompt_token_ids", [1]))
        assert o.outputs[0].token_ids == dec[n0:], (o.outputs[0].token_ids, dec[n0:])
    llm.shutdown()


@pytest.mark.parametrize("gen", ["qwen2vl-tiny", "qwen25vl-tiny"])
def test_qwen2_vl_matches_hf(gen):
    """Dynamic-resolution ViT + M-RoPE decoder: image patches, grid, 3-D position ids and the decode position delta
    reproduce the HF model's greedy continuation for image+text and for text-only prompts in one batch."""
    import numpy as np
    import transformers
    from PIL import Image
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config(gen)
    is25 = gen == "qwen25vl-tiny"
    text_keys = ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                 "num_key_value_heads", "max_position_embeddings", "rms_norm_eps", "hidden_act", "tie_word_embeddings")
    tc = {k: cfg[k] for k in text_keys}
    tc["rope_parameters"] = {"rope_type": "default", "rope_theta": cfg["rope_theta"], "mrope_section": [2, 3, 3]}
    CfgCls = transformers.Qwen2_5_VLConfig if is25 else transformers.Qwen2VLConfig
    ModelCls = transformers.Qwen2_5_VLForConditionalGeneration if is25 else transformers.Qwen2VLForConditionalGeneration
    hf_cfg = CfgCls(text_config=tc, vision_config=cfg["vision_config"], image_token_id=500,
                    video_token_id=501, vision_start_token_id=502, vision_end_token_id=503)
    torch.manual_seed(0)
    hf = ModelCls(hf_cfg).float().eval()
    rng = np.random.default_rng(0)
    img = Image.fromarray(rng.integers(0, 255, (84, 140, 3) if is25 else (56, 84, 3), dtype=np.uint8))
    ip = transformers.Qwen2VLImageProcessor()
    px = ip(images=[img], return_tensors="pt")
    n_img = int(px["image_grid_thw"][0].prod()) // 4
    pre, post = [5, 17, 99, 502], [503, 3, 250, 41, 7]
    ids_hf = pre + [500] * n_img + post
    with torch.no_grad():
        ref_img = hf.generate(input_ids=torch.tensor([ids_hf]), attention_mask=torch.ones(1, len(ids_hf), dtype=torch.long),
                              pixel_values=px["pixel_values"], image_grid_thw=px["image_grid_thw"], max_new_tokens=8,
                              do_sample=False, pad_token_id=0, eos_token_id=None)[0, len(ids_hf):].tolist()
    text_prompt = [5, 17, 99, 3, 250, 41, 5, 17, 99, 3]
    ref_txt = hf_greedy(hf, text_prompt, 8)
    llm = make_llm(gen, weights=hf.state_dict().items(), enable_prefix_caching=False)
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    outs = llm.generate([{"prompt_token_ids": pre + [500] + post, "multi_modal_data": {"image": img}},
                         {"prompt_token_ids": text_prompt}], sp)
    llm.shutdown()
    assert outs[0].outputs[0].token_ids == ref_img, (outs[0].outputs[0].token_ids, ref_img)
    assert outs[1].outputs[0].token_ids == ref_txt


@pytest.mark.parametrize("gen", ["qwen2vl-tiny", "qwen25vl-tiny"])
def test_qwen2_vl_video_matches_hf(gen):
    """Video input: frames are grouped into temporal patches, each temporal patch attends within itself in the ViT, and
    the placeholders get (t, h, w) M-RoPE ids (Qwen2.5-VL: t stretched by seconds-per-patch x tokens_per_second)."""
    import numpy as np
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config(gen)
    is25 = gen == "qwen25vl-tiny"
    text_keys = ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                 "num_key_value_heads", "max_position_embeddings", "rms_norm_eps", "hidden_act", "tie_word_embeddings")
    tc = {k: cfg[k] for k in text_keys}
    tc["rope_parameters"] = {"rope_type": "default", "rope_theta": cfg["rope_theta"], "mrope_section": [2, 3, 3]}
    CfgCls = transformers.Qwen2_5_VLConfig if is25 else transformers.Qwen2VLConfig
    ModelCls = transformers.Qwen2_5_VLForConditionalGeneration if is25 else transformers.Qwen2VLForConditionalGeneration
    hf_cfg = CfgCls(text_config=tc, vision_config=cfg["vision_config"], image_token_id=500,
                    video_token_id=501, vision_start_token_id=502, vision_end_token_id=503)
    torch.manual_seed(0)
    hf = ModelCls(hf_cfg).float().eval()
    rng = np.random.default_rng(1)
    vid = rng.integers(0, 255, (6, 84, 112, 3) if is25 else (4, 56, 84, 3), dtype=np.uint8)
    vp = transformers.Qwen2VLVideoProcessor(size={"shortest_edge": 56 * 56, "longest_edge": 28 * 28 * 1280})
    px = vp(videos=[torch.from_numpy(vid).permute(0, 3, 1, 2)], return_tensors="pt", do_sample_frames=False)
    n_vid = int(px["video_grid_thw"][0].prod()) // 4
    pre, post = [5, 17, 99, 502], [503, 3, 250, 41, 7]
    ids_hf = pre + [501] * n_vid + post
    kw = {"second_per_grid_ts": torch.tensor([0.5])} if is25 else {}
    with torch.no_grad():
        ref = hf.generate(input_ids=torch.tensor([ids_hf]), attention_mask=torch.ones(1, len(ids_hf), dtype=torch.long),
                          pixel_values_videos=px["pixel_values_videos"], video_grid_thw=px["video_grid_thw"],
                          max_new_tokens=8, do_sample=False, pad_token_id=0, eos_token_id=None, **kw)[0, len(ids_hf):].tolist()
    llm = make_llm(gen, weights=hf.state_dict().items(), enable_prefix_caching=False,
                   hf_overrides={"video_token_id": 501})
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    mm = {"video": vid, "fps": 4.0} if is25 else {"video": [f for f in vid]}      # 2 frames / 4 fps = 0.5 s per patch
    outs = llm.generate([{"prompt_token_ids": pre + [501] + post, "multi_modal_data": mm}], sp)
    llm.shutdown()
    assert outs[0].outputs[0].token_ids == ref, (outs[0].outputs[0].token_ids, ref)


def test_decoder_pooling_adapters_match_hf():
    """A registered decoder as (a) sequence classifier: last-token state → score head → softmax == HF
    LlamaForSequenceClassification; (b) embedding model loaded from a bare ``LlamaModel`` checkpoint."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    kw = {k: v for k, v in cfg.items() if k not in ("architectures", "model_type", "torch_dtype")}
    hf_cfg = transformers.AutoConfig.for_model("llama", **kw, num_labels=3)
    hf_cfg.pad_token_id = 0
    torch.manual_seed(0)
    hf = transformers.LlamaForSequenceClassification(hf_cfg).float().eval()
    prompts = [[5, 17, 99, 3, 250, 41], list(range(10, 45))]
    llm = make_llm("llama-tiny", hf_overrides={"architectures": ["LlamaForSequenceClassification"], "num_labels": 3},
                   runner="pooling")
    runner = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner
    assert type(runner.model).__name__ == "LlamaForSequenceClassification"
    runner.model.load_weights(hf.state_dict().items())
    outs = llm.classify([{"prompt_token_ids": p} for p in prompts])
    for p, o in zip(prompts, outs):
        with torch.no_grad():
            ref = hf(torch.tensor([p])).logits[0].softmax(-1)
        torch.testing.assert_close(torch.as_tensor(o.outputs.data).float(), ref, atol=2e-4, rtol=1e-3)
    llm.shutdown()
    torch.manual_seed(1)
    base = transformers.LlamaModel(hf_cfg).float().eval()
    llm = make_llm("llama-tiny", hf_overrides={"architectures": ["LlamaModel"]}, runner="pooling")
    llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model.load_weights(base.state_dict().items())
    outs = llm.embed([{"prompt_token_ids": p} for p in prompts])
    for p, o in zip(prompts, outs):
        with torch.no_grad():
            ref = torch.nn.functional.normalize(base(torch.tensor([p])).last_hidden_state[0, -1], dim=-1)
        torch.testing.assert_close(torch.as_tensor(o.outputs.data).float(), ref, atol=2e-4, rtol=1e-3)
    llm.shutdown()


def test_modernbert_embeddings_match_hf():
    """Rotary encoder with alternating global / symmetric-local attention through the pooling runner == HF."""
    hf = make_hf_model("modernbert-tiny")
    hf.config._attn_implementation = "eager"
    llm = make_llm("modernbert-tiny", weights=hf.state_dict().items(), enable_prefix_caching=False)
    prompts = [[3, 17, 99, 250, 41, 7], list(range(10, 55))]
    outs = llm.embed([{"prompt_token_ids": p} for p in prompts])
    for p, o in zip(prompts, outs):
        with torch.no_grad():
            ref = hf(torch.tensor([p])).last_hidden_state[0, 0]
        ref = torch.nn.functional.normalize(ref, dim=-1)
        torch.testing.assert_close(torch.as_tensor(o.outputs.data).float(), ref, atol=2e-4, rtol=1e-3)
    llm.shutdown()


def test_dbrx_matches_hf():
    """DBRX against transformers.  transformers >= 5 keeps each expert matrix transposed relative to the released
    checkpoints ([D, I] slabs under swapped config names); the test converts to the released [E·I, D] layout, which is
    what the loader accepts."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("dbrx-tiny")
    D, I, E = cfg["d_model"], cfg["ffn_config"]["ffn_hidden_size"], cfg["ffn_config"]["moe_num_experts"]
    ffn = dict(cfg["ffn_config"], ffn_hidden_size=D, hidden_size=I)
    hf_cfg = transformers.DbrxConfig(d_model=D, n_heads=cfg["n_heads"], n_layers=cfg["n_layers"], max_seq_len=cfg["max_seq_len"],
                                     vocab_size=cfg["vocab_size"], attn_config=cfg["attn_config"], ffn_config=ffn,
                                     tie_word_embeddings=False)
    torch.manual_seed(0)
    hf = transformers.DbrxForCausalLM(hf_cfg).float().eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():      # (the default init leaves the expert slabs near zero)
            if ".experts.mlp." in n or "Wqkv" in n:
                p.copy_(torch.randn_like(p) * 0.3)
    sd = {}
    for k, v in hf.state_dict().items():
        if ".experts.mlp." in k:
            v = v.view(E, D, I).transpose(1, 2).reshape(E * I, D).contiguous()
        sd[k] = v
    llm = make_llm("dbrx-tiny", weights=sd.items())
    sp = SamplingParams(max_tokens=10, temperature=0.0, ignore_eos=True)
    outs = llm.generate([{"prompt_token_ids": p} for p in PROMPTS], sp)
    for p, o in zip(PROMPTS, outs):
        assert o.outputs[0].token_ids == hf_greedy(hf, p, 10)


def _gen(preset, weights, overrides=None, n=8):
    llm = make_llm(preset, weights=weights, **({"hf_overrides": overrides} if overrides else {}))
    sp = SamplingParams(max_tokens=n, temperature=0.0, ignore_eos=True, logprobs=1)
    outs = llm.generate([{"prompt_token_ids": p} for p in PROMPTS], sp)
    llm.shutdown()
    return [(list(o.outputs[0].token_ids), [lp[t].logprob for lp, t in zip(o.outputs[0].logprobs, o.outputs[0].token_ids)])
            for o in outs]


def _same(a, b, tol=1e-4):
    for (ta, la), (tb, lb) in zip(a, b):
        assert ta == tb
        assert max(abs(x - y) for x, y in zip(la, lb)) < tol


def test_llama_aliases_match_llama_on_converted_checkpoints():
    """InternLM2 (packed wqkv, w1/w3/w2), EXAONE-3 (GPT-style names), InternLM3 / IQuest-Coder / TeleChat3 (aliases), MiMo
    (Qwen2 + ignored MTP heads): the same weights under each family's checkpoint naming give the Llama outputs."""
    from tests.utils import llama_full_weights
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    w = dict(llama_full_weights(cfg, seed=3))
    ref = _gen("llama-tiny", list(w.items()))
    nh, nkv = cfg["num_attention_heads"], cfg.get("num_key_value_heads") or cfg["num_attention_heads"]
    hd, g = cfg["hidden_size"] // nh, nh // nkv

    # ---- InternLM2
    il = {}
    for k, v in w.items():
        if ".self_attn.q_proj." in k:
            q, kk, vv = v, w[k.replace("q_proj", "k_proj")], w[k.replace("q_proj", "v_proj")]
            H = q.shape[1]
            packed = torch.cat([q.view(nkv, g, hd, H), kk.view(nkv, 1, hd, H), vv.view(nkv, 1, hd, H)], dim=1)
            il[k.replace(".self_attn.q_proj.", ".attention.wqkv.")] = packed.reshape(-1, H)
        elif ".self_attn.k_proj." in k or ".self_attn.v_proj." in k:
            continue
        else:
            k2 = (k.replace("model.embed_tokens.", "model.tok_embeddings.").replace("lm_head.weight", "output.weight")
                  .replace(".input_layernorm.", ".attention_norm.").replace(".post_attention_layernorm.", ".ffn_norm.")
                  .replace(".self_attn.o_proj.", ".attention.wo.").replace(".mlp.gate_proj.", ".feed_forward.w1.")
                  .replace(".mlp.up_proj.", ".feed_forward.w3.").replace(".mlp.down_proj.", ".feed_forward.w2."))
            il[k2] = v
    _same(ref, _gen("llama-tiny", list(il.items()), {"architectures": ["InternLM2ForCausalLM"]}))

    # ---- EXAONE-3
    ex = {}
    for k, v in w.items():
        k2 = (k.replace("model.embed_tokens.", "transformer.wte.").replace("model.norm.", "transformer.ln_f.")
              .replace("model.layers.", "transformer.h.").replace(".self_attn.o_proj.", ".attn.attention.out_proj.")
              .replace(".self_attn.", ".attn.attention.").replace(".input_layernorm.", ".ln_1.")
              .replace(".post_attention_layernorm.", ".ln_2.").replace(".mlp.gate_proj.", ".mlp.c_fc_0.")
              .replace(".mlp.up_proj.", ".mlp.c_fc_1.").replace(".mlp.down_proj.", ".mlp.c_proj."))
        ex[k2] = v
    ov = {"architectures": ["ExaoneForCausalLM"], "num_layers": cfg["num_hidden_layers"], "layer_norm_epsilon": cfg.get("rms_norm_eps", 1e-5),
          "activation_function": "silu"}
    _same(ref, _gen("llama-tiny", list(ex.items()), ov))

    # ---- plain aliases
    for arch in ("InternLM3ForCausalLM", "IQuestCoderForCausalLM", "TeleChat3ForCausalLM", "TransformersForCausalLM"):
        _same(ref, _gen("llama-tiny", list(w.items()), {"architectures": [arch]}), tol=2e-4)


def test_mup_llamas_match_granite_parameterisation():
    """MiniCPM's scale_emb / scale_depth / dim_model_base and TeleFLM's input / output multipliers are the Granite
    multipliers under other names: identical weights, identical outputs."""
    from tests.utils import llama_full_weights
    from vllm_b200.config import get_preset_hf_config
    import math
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=5))
    L, H = cfg["num_hidden_layers"], cfg["hidden_size"]
    hd = H // cfg["num_attention_heads"]
    gr = {"architectures": ["GraniteForCausalLM"], "embedding_multiplier": 6.0, "residual_multiplier": 1.4 / math.sqrt(L),
          "logits_scaling": H / 32.0, "attention_multiplier": hd ** -0.5}
    mc = {"architectures": ["MiniCPMForCausalLM"], "scale_emb": 6.0, "scale_depth": 1.4, "dim_model_base": 32}
    _same(_gen("llama-tiny", w, gr), _gen("llama-tiny", w, mc))
    gr2 = {"architectures": ["GraniteForCausalLM"], "embedding_multiplier": 3.0, "residual_multiplier": 1.0, "logits_scaling": 4.0,
           "attention_multiplier": hd ** -0.5}
    tf = {"architectures": ["TeleFLMForCausalLM"], "use_mup": True, "input_mult": 3.0, "output_mult": 2.0, "mup_scale_factor": 8.0}
    _same(_gen("llama-tiny", w, gr2), _gen("llama-tiny", w, tf))
    # MiMo: Qwen2 with extra multi-token-prediction tensors in the checkpoint
    q = _gen("llama-tiny", w, {"architectures": ["Qwen2ForCausalLM"], "attention_bias": False})
    extra = w + [("model.mtp_layers.0.input_layernorm.weight", torch.ones(H))]
    _same(q, _gen("llama-tiny", extra, {"architectures": ["MiMoForCausalLM"], "attention_bias": False}))


def test_deepseek_v1_moe_and_orion_run():
    """Families without a transformers implementation to compare against: DeepSeek-MoE v1 (softmax router, shared experts,
    full rotary) must equal the GLM-4.5 composition configured the same way; Orion (LayerNorm Llama) must generate."""
    from vllm_b200.config import get_preset_hf_config
    base = {"use_qk_norm": False, "partial_rotary_factor": 1.0, "scoring_func": "softmax", "topk_method": "greedy",
            "norm_topk_prob": False, "n_group": 1, "topk_group": 1}
    a = _gen("dots1-tiny", None, dict(base, architectures=["Glm4MoeForCausalLM"]))
    b = _gen("dots1-tiny", None, {"architectures": ["DeepseekForCausalLM"], "use_qk_norm": False, "n_group": 1, "topk_group": 1})
    _same(a, b)
    o = _gen("llama-tiny", None, {"architectures": ["OrionForCausalLM"]})
    assert all(len(t) == 8 for t, _ in o)


def _hf_from_preset(preset, cls_name, model_type=None, **extra):
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config(preset)
    kw = {k: v for k, v in cfg.items() if k not in ("architectures", "model_type", "torch_dtype")}
    kw.update(extra)
    hf_cfg = transformers.AutoConfig.for_model(model_type or cfg["model_type"], **kw)
    torch.manual_seed(0)
    return getattr(transformers, cls_name)(hf_cfg).float().eval()


def _pooling_llm(preset, arch, hf, **ov):
    llm = make_llm(preset, hf_overrides=dict(architectures=[arch], **ov), runner="pooling", enable_prefix_caching=False)
    m = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model
    m.load_weights(hf.state_dict().items())
    return llm, m


def test_classification_heads_match_hf():
    """Sequence and token classification heads of the encoder families and a decoder with an irregular base name:
    RoBERTa (dense-tanh-out_proj on <s>), BERT token classifier, ModernBERT prediction head (CLS and per token), GPT-2
    last-token classifier — probabilities equal to transformers."""
    prompts = [[3, 17, 99, 250, 41, 7], list(range(10, 45))]
    # RoBERTa: position ids start at padding_idx + 1 → compare with explicit position ids
    hf = _hf_from_preset("bert-tiny", "RobertaForSequenceClassification", "roberta", num_labels=3, pad_token_id=1,
                         max_position_embeddings=258)
    llm, m = _pooling_llm("bert-tiny", "RobertaForSequenceClassification", hf, num_labels=3, model_type="roberta",
                          max_position_embeddings=258)
    assert type(m).__name__ == "RobertaForSequenceClassification"
    for p, o in zip(prompts, llm.classify([{"prompt_token_ids": p} for p in prompts])):
        with torch.no_grad():
            ref = hf(torch.tensor([p]), position_ids=torch.arange(2, 2 + len(p))[None]).logits[0].softmax(-1)
        torch.testing.assert_close(torch.as_tensor(o.outputs.data).float(), ref, atol=2e-4, rtol=1e-3)
    llm.shutdown()
    # BERT token classification
    hf = _hf_from_preset("bert-tiny", "BertForTokenClassification", num_labels=5)
    llm, m = _pooling_llm("bert-tiny", "BertForTokenClassification", hf, num_labels=5)
    outs = llm.encode([{"prompt_token_ids": p} for p in prompts], pooling_task="token_classify")
    for p, o in zip(prompts, outs):
        with torch.no_grad():
            ref = hf(torch.tensor([p])).logits[0].softmax(-1)
        got = torch.as_tensor(o.outputs.data).float()
        assert got.shape == (len(p), 5)
        torch.testing.assert_close(got, ref, atol=2e-4, rtol=1e-3)
    llm.shutdown()
    # ModernBERT sequence + token heads
    for cls, task in (("ModernBertForSequenceClassification", "classify"), ("ModernBertForTokenClassification", "token_classify")):
        hf = _hf_from_preset("modernbert-tiny", cls, num_labels=4, classifier_bias=True)
        llm, m = _pooling_llm("modernbert-tiny", cls, hf, num_labels=4, classifier_bias=True)
        outs = llm.encode([{"prompt_token_ids": p} for p in prompts], pooling_task=task)
        for p, o in zip(prompts, outs):
            with torch.no_grad():
                ref = hf(torch.tensor([p])).logits[0].softmax(-1)
            torch.testing.assert_close(torch.as_tensor(o.outputs.data).float().reshape(ref.shape), ref, atol=3e-4, rtol=1e-3)
        llm.shutdown()
    # GPT-2 (causal base class is GPT2LMHeadModel, not GPT2ForCausalLM)
    hf = _hf_from_preset("gpt2-tiny", "GPT2ForSequenceClassification", num_labels=3, pad_token_id=0)
    llm, m = _pooling_llm("gpt2-tiny", "GPT2ForSequenceClassification", hf, num_labels=3)
    for p, o in zip(prompts, llm.classify([{"prompt_token_ids": p} for p in prompts])):
        with torch.no_grad():
            ref = hf(torch.tensor([p])).logits[0].softmax(-1)
        torch.testing.assert_close(torch.as_tensor(o.outputs.data).float(), ref, atol=2e-4, rtol=1e-3)
    llm.shutdown()


def test_reward_model_heads():
    """Per-token reward heads: Qwen2 PRM (Linear-ReLU-Linear → 2 classes per position) and the InternLM2 ``v_head``."""
    from tests.utils import llama_full_weights
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    H = cfg["hidden_size"]
    w = list(llama_full_weights(cfg, seed=2))
    g = torch.Generator().manual_seed(0)
    head = {"score.0.weight": torch.randn(H, H, generator=g) * 0.1, "score.0.bias": torch.randn(H, generator=g) * 0.1,
            "score.2.weight": torch.randn(2, H, generator=g) * 0.1, "score.2.bias": torch.randn(2, generator=g) * 0.1}
    prompts = [[5, 17, 99, 3, 250, 41]]
    llm = make_llm("llama-tiny", hf_overrides={"architectures": ["Qwen2ForProcessRewardModel"], "attention_bias": False},
                   runner="pooling")
    m = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model
    m.load_weights(w + list(head.items()))
    out = llm.encode([{"prompt_token_ids": p} for p in prompts], pooling_task="token_classify")[0]
    got = torch.as_tensor(out.outputs.data).float()
    llm.shutdown()
    # reference: hidden states of the embedding adapter of the same weights, head applied by hand
    llm = make_llm("llama-tiny", hf_overrides={"architectures": ["Qwen2Model"], "attention_bias": False}, runner="pooling")
    llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model.load_weights(w)
    from vllm_b200 import PoolingParams
    hid = torch.as_tensor(llm.encode([{"prompt_token_ids": prompts[0]}], PoolingParams(task="token_embed", normalize=False))[0].outputs.data).float()
    llm.shutdown()
    ref = (torch.relu(hid @ head["score.0.weight"].t() + head["score.0.bias"]) @ head["score.2.weight"].t() + head["score.2.bias"]).softmax(-1)
    assert got.shape == (6, 2)
    torch.testing.assert_close(got, ref, atol=2e-4, rtol=1e-3)
    from vllm_b200.models.registry import resolve_model_cls
    assert resolve_model_cls("InternLM2ForRewardModel").__name__ == "InternLM2RewardModel"
    assert resolve_model_cls("Gemma3TextModel").__name__.startswith("Gemma3")


def _vlm_text_cfg(cfg, drop):
    return {k: v for k, v in cfg.items() if k not in drop}


def test_smolvlm_matches_hf_with_image():
    """SigLIP tower + pixel-shuffle connector + Llama decoder == HF SmolVLMForConditionalGeneration (greedy), and the same
    weights under the Idefics3 architecture name."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("smolvlm-tiny")
    text = _vlm_text_cfg(cfg, ("architectures", "model_type", "vision_config", "image_token_id", "scale_factor", "torch_dtype"))
    hf_cfg = transformers.SmolVLMConfig(vision_config=cfg["vision_config"], text_config=dict(model_type="llama", **text),
                                        scale_factor=2, image_token_id=511, tie_word_embeddings=False)
    torch.manual_seed(0)
    hf = transformers.SmolVLMForConditionalGeneration(hf_cfg).eval()
    g = torch.Generator().manual_seed(1)
    img = torch.randn(3, 32, 32, generator=g)
    ids = [5, 17, 511, 99, 3, 250]                       # one <image> placeholder → (32/8)^2 / 4 = 4 tokens
    exp_ids = [5, 17] + [511] * 4 + [99, 3, 250]
    with torch.no_grad():
        ref = hf.generate(input_ids=torch.tensor([exp_ids]), pixel_values=img[None, None], max_new_tokens=8, do_sample=False,
                          pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    for arch in ("SmolVLMForConditionalGeneration", "Idefics3ForConditionalGeneration"):
        llm = make_llm("smolvlm-tiny", weights=hf.state_dict().items(), hf_overrides={"architectures": [arch]})
        out = llm.generate([{"prompt_token_ids": ids, "multi_modal_data": {"image": img}}, {"prompt_token_ids": [7, 8, 9, 10]}], sp)
        assert out[0].outputs[0].token_ids == ref
        llm.shutdown()


def test_gemma3_multimodal_matches_hf_with_image():
    """SigLIP tower → pooled, normed, projected image tokens → Gemma-3 decoder == HF Gemma3ForConditionalGeneration (greedy;
    HF without ``token_type_ids`` uses the causal mask as well)."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("gemma3mm-tiny")
    text = _vlm_text_cfg(cfg, ("architectures", "model_type", "vision_config", "image_token_index", "mm_tokens_per_image", "torch_dtype"))
    hf_cfg = transformers.Gemma3Config(vision_config=cfg["vision_config"], text_config=dict(model_type="gemma3_text", **text),
                                       mm_tokens_per_image=4, image_token_index=511, boi_token_index=509, eoi_token_index=510)
    torch.manual_seed(0)
    hf = transformers.Gemma3ForConditionalGeneration(hf_cfg).eval()
    with torch.no_grad():
        hf.model.multi_modal_projector.mm_input_projection_weight.normal_(0, 0.1)
    g = torch.Generator().manual_seed(1)
    img = torch.randn(3, 32, 32, generator=g)
    ids = [5, 17, 511, 99, 3, 250]
    exp_ids = [5, 17] + [511] * 4 + [99, 3, 250]
    with torch.no_grad():
        ref = hf.generate(input_ids=torch.tensor([exp_ids]), pixel_values=img[None], max_new_tokens=8, do_sample=False,
                          pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
    llm = make_llm("gemma3mm-tiny", weights=hf.state_dict().items())
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    out = llm.generate([{"prompt_token_ids": ids, "multi_modal_data": {"image": img}}], sp)
    assert out[0].outputs[0].token_ids == ref
    llm.shutdown()


def test_cohere_vision_models_match_hf_with_image():
    """Command-A Vision and Aya Vision: SigLIP → pixel shuffle → SwiGLU projector (with / without the leading LayerNorm,
    post-norm / pre-norm tower features) → Cohere-2 decoder, against transformers."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("cohere2vision-tiny")
    drop = ("architectures", "model_type", "vision_config", "image_token_id", "downsample_factor", "alignment_intermediate_size", "torch_dtype")
    text = dict(model_type="cohere2", **_vlm_text_cfg(cfg, drop))
    g = torch.Generator().manual_seed(1)
    img = torch.randn(3, 32, 32, generator=g)
    ids = [5, 17, 511, 99, 3, 250]
    exp_ids = [5, 17] + [511] * 4 + [99, 3, 250]
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    for arch, cfg_cls, kw in (("Cohere2VisionForConditionalGeneration", "Cohere2VisionConfig", {}),
                              ("AyaVisionForConditionalGeneration", "AyaVisionConfig",
                               dict(vision_feature_layer=-1, vision_feature_select_strategy="full", adapter_layer_norm_eps=1e-6))):
        hf_cfg = getattr(transformers, cfg_cls)(vision_config=cfg["vision_config"], text_config=text, downsample_factor=2,
                                                alignment_intermediate_size=96, image_token_id=511, **kw)
        if arch.startswith("Aya"):
            hf_cfg.image_token_index = 511
        torch.manual_seed(0)
        hf = getattr(transformers, arch)(hf_cfg).eval()
        with torch.no_grad():
            ref = hf.generate(input_ids=torch.tensor([exp_ids]), pixel_values=img[None], max_new_tokens=8, do_sample=False,
                              pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
        llm = make_llm("cohere2vision-tiny", weights=hf.state_dict().items(), hf_overrides={"architectures": [arch]})
        out = llm.generate([{"prompt_token_ids": ids, "multi_modal_data": {"image": img}}], sp)
        assert out[0].outputs[0].token_ids == ref, arch
        llm.shutdown()


def test_internvl_matches_hf_with_image():
    """InternViT (class token, layer scales, q/k RMSNorm) → pixel shuffle → LN-MLP projector → Qwen2 decoder == HF
    InternVLForConditionalGeneration (greedy)."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("internvl-tiny")
    drop = ("architectures", "model_type", "vision_config", "image_token_id", "downsample_ratio", "vision_feature_layer",
            "vision_feature_select_strategy", "projector_hidden_act", "text_model_type", "torch_dtype")
    hf_cfg = transformers.InternVLConfig(vision_config=cfg["vision_config"], text_config=dict(model_type="qwen2", **_vlm_text_cfg(cfg, drop)),
                                         image_token_id=511, downsample_ratio=0.5)
    torch.manual_seed(0)
    hf = transformers.InternVLForConditionalGeneration(hf_cfg).eval()
    g = torch.Generator().manual_seed(1)
    img = torch.randn(3, 32, 32, generator=g)
    ids = [5, 17, 511, 99, 3, 250]
    exp_ids = [5, 17] + [511] * 4 + [99, 3, 250]
    with torch.no_grad():
        ref = hf.generate(input_ids=torch.tensor([exp_ids]), pixel_values=img[None], max_new_tokens=8, do_sample=False,
                          pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
    llm = make_llm("internvl-tiny", weights=hf.state_dict().items())
    out = llm.generate([{"prompt_token_ids": ids, "multi_modal_data": {"image": img}}],
                       SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True))
    assert out[0].outputs[0].token_ids == ref
    llm.shutdown()


def test_audio_language_models_match_hf():
    """Qwen2-Audio (masked, pooled Whisper tower + linear projector; a full window and a clip shorter than the window) and
    Voxtral (tower + frame-group GELU projector) against transformers, greedy, with precomputed log-mel inputs."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    g = torch.Generator().manual_seed(1)
    # ---- Qwen2-Audio
    cfg = get_preset_hf_config("qwen2audio-tiny")
    drop = ("architectures", "model_type", "audio_config", "audio_token_index", "text_model_type", "torch_dtype")
    hf_cfg = transformers.Qwen2AudioConfig(audio_config=cfg["audio_config"], text_config=dict(model_type="qwen2", **_vlm_text_cfg(cfg, drop)),
                                           audio_token_index=511)
    torch.manual_seed(0)
    hf = transformers.Qwen2AudioForConditionalGeneration(hf_cfg).eval()
    llm = make_llm("qwen2audio-tiny", weights=hf.state_dict().items())
    for valid in (48, 30):
        mel = torch.randn(16, 48, generator=g)
        mel[:, valid:] = -1.0
        n = ((valid - 1) // 2 + 1 - 2) // 2 + 1
        ids = [5, 17, 511, 99, 3, 250]
        exp_ids = [5, 17] + [511] * n + [99, 3, 250]
        fam = torch.zeros(1, 48, dtype=torch.long)
        fam[:, :valid] = 1
        with torch.no_grad():
            ref = hf.generate(input_ids=torch.tensor([exp_ids]), input_features=mel[None], feature_attention_mask=fam,
                              attention_mask=torch.ones(1, len(exp_ids), dtype=torch.long), max_new_tokens=8, do_sample=False,
                              pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
        out = llm.generate([{"prompt_token_ids": ids, "multi_modal_data": {"audio": mel[:, :valid]}}], sp)
        assert out[0].outputs[0].token_ids == ref, valid
    llm.shutdown()
    # ---- Voxtral
    cfg = get_preset_hf_config("voxtral-tiny")
    drop = ("architectures", "model_type", "audio_config", "audio_token_id", "text_model_type", "projector_hidden_act", "torch_dtype")
    hf_cfg = transformers.VoxtralConfig(audio_config=cfg["audio_config"], text_config=dict(model_type="llama", **_vlm_text_cfg(cfg, drop)),
                                        audio_token_id=511)
    torch.manual_seed(0)
    hf = transformers.VoxtralForConditionalGeneration(hf_cfg).eval()
    mel = torch.randn(16, 48, generator=g)
    n = 24 * 32 // 64
    ids = [5, 17, 511, 99, 3, 250]
    exp_ids = [5, 17] + [511] * n + [99, 3, 250]
    with torch.no_grad():
        ref = hf.generate(input_ids=torch.tensor([exp_ids]), input_features=mel[None], max_new_tokens=8, do_sample=False,
                          pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
    llm = make_llm("voxtral-tiny", weights=hf.state_dict().items())
    out = llm.generate([{"prompt_token_ids": ids, "multi_modal_data": {"audio": mel}}], sp)
    assert out[0].outputs[0].token_ids == ref
    llm.shutdown()


def test_mistral3_matches_hf_with_image():
    """Pixtral tower (2-D rotary) → norm → learned 2×2 patch merger → projector → Mistral decoder == HF
    Mistral3ForConditionalGeneration, with the [IMG]…[IMG_BREAK]…[IMG_END] row layout produced by the processor."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("mistral3-tiny")
    drop = ("architectures", "model_type", "vision_config", "image_token_index", "spatial_merge_size", "multimodal_projector_bias",
            "projector_hidden_act", "vision_feature_layer", "text_model_type", "torch_dtype")
    hf_cfg = transformers.Mistral3Config(vision_config=cfg["vision_config"], text_config=dict(model_type="mistral", **_vlm_text_cfg(cfg, drop)),
                                         image_token_index=10, spatial_merge_size=2, multimodal_projector_bias=False, vision_feature_layer=-1)
    torch.manual_seed(0)
    hf = transformers.Mistral3ForConditionalGeneration(hf_cfg).eval()
    g = torch.Generator().manual_seed(1)
    img = torch.randn(3, 32, 32, generator=g)
    ids = [5, 17, 10, 99, 3, 250]
    rows = []
    for r in range(4):                                     # 32 / 4 / 2 = 4 merged patches per side
        rows += [10] * 4 + [13 if r == 3 else 12]
    exp_ids = [5, 17] + rows + [99, 3, 250]
    with torch.no_grad():
        ref = hf.generate(input_ids=torch.tensor([exp_ids]), pixel_values=img[None], image_sizes=torch.tensor([[32, 32]]),
                          max_new_tokens=8, do_sample=False, pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
    llm = make_llm("mistral3-tiny", weights=hf.state_dict().items())
    out = llm.generate([{"prompt_token_ids": ids, "multi_modal_data": {"image": img}}],
                       SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True))
    assert out[0].prompt_token_ids == exp_ids
    assert out[0].outputs[0].token_ids == ref
    llm.shutdown()


def test_llava_next_anyres_matches_hf():
    """Any-resolution LLaVA-NeXT: a wide image on a 1×2 tile grid (padding rows cropped, newline token per feature row) and
    the same model on a square single-canvas image, against transformers; plus the built-in tiler's token count."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = dict(get_preset_hf_config("llava-tiny"))
    pins = [[28, 28], [28, 56], [56, 28], [56, 56]]
    drop = ("architectures", "model_type", "vision_config", "image_token_index", "vision_feature_layer",
            "vision_feature_select_strategy", "projector_hidden_act", "torch_dtype")
    hf_cfg = transformers.LlavaNextConfig(vision_config=cfg["vision_config"], text_config=dict(model_type="llama", **_vlm_text_cfg(cfg, drop)),
                                          image_token_index=511, vision_feature_layer=-2, vision_feature_select_strategy="default",
                                          projector_hidden_act="gelu", image_grid_pinpoints=pins)
    torch.manual_seed(0)
    hf = transformers.LlavaNextForConditionalGeneration(hf_cfg).eval()
    with torch.no_grad():
        hf.model.image_newline.normal_(0, 0.5)
    llm = make_llm("llava-tiny", weights=hf.state_dict().items(),
                   hf_overrides={"architectures": ["LlavaNextForConditionalGeneration"], "image_grid_pinpoints": pins})
    proc = llm.llm_engine.processor.mm_processor if hasattr(llm.llm_engine, "processor") else None
    g = torch.Generator().manual_seed(1)
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    for size, n_tiles in (((20, 50), 3), ((40, 44), 5)):
        tiles = torch.randn(n_tiles, 3, 28, 28, generator=g)
        with torch.no_grad():
            feats = hf.get_image_features(pixel_values=tiles[None], image_sizes=torch.tensor([size]),
                                          vision_feature_layer=-2, vision_feature_select_strategy="default").pooler_output
        n = feats[0].shape[0] if isinstance(feats, (list, tuple)) else feats.shape[0]
        exp_ids = [5, 17] + [511] * n + [99, 3, 250]
        with torch.no_grad():
            ref = hf.generate(input_ids=torch.tensor([exp_ids]), pixel_values=tiles[None], image_sizes=torch.tensor([size]),
                              max_new_tokens=8, do_sample=False, pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
        out = llm.generate([{"prompt_token_ids": [5, 17, 511, 99, 3, 250],
                             "multi_modal_data": {"image": {"tiles": tiles, "image_size": size}}}], sp)
        assert len(out[0].prompt_token_ids) == len(exp_ids), (size, n)
        assert out[0].outputs[0].token_ids == ref, size
    # the built-in tiler: raw HWC uint8 image → canvas fit + tiles, token count consistent with the packer
    from vllm_b200.models.llava_next import LlavaNextProcessor
    import numpy as np
    mc = llm.llm_engine.vllm_config.model_config
    p = LlavaNextProcessor(mc)
    img = (np.random.RandomState(0).rand(20, 50, 3) * 255).astype(np.uint8)
    ids, feats = p.process([5, 511, 7], {"image": img})
    assert feats[0]["data"].shape == (3, 3, 28, 28) and feats[0]["image_size"] == (20, 50)
    assert len(ids) == 2 + feats[0]["length"]
    out = llm.generate([{"prompt_token_ids": [5, 511, 7], "multi_modal_data": {"image": img}}], sp)
    assert len(out[0].outputs[0].token_ids) == 8
    llm.shutdown()


def test_nomic_bert_embeddings_match_hf():
    """Rotary, SwiGLU, post-norm encoder: mean-pooled, normalised embeddings == HF NomicBertModel."""
    hf = make_hf_model("nomicbert-tiny")
    llm = make_llm("nomicbert-tiny", weights=hf.state_dict().items())
    prompts = [[3, 17, 99, 250, 41, 7], list(range(10, 45))]
    outs = llm.embed([{"prompt_token_ids": p} for p in prompts])
    for p, o in zip(prompts, outs):
        with torch.no_grad():
            ref = torch.nn.functional.normalize(hf(torch.tensor([p])).last_hidden_state[0].mean(0), dim=-1)
        torch.testing.assert_close(torch.as_tensor(o.outputs.data).float(), ref, atol=2e-4, rtol=1e-3)
    llm.shutdown()


def test_colbert_and_splade_heads_on_bert():
    """Late-interaction (per-token projected, normalised vectors) and SPLADE (max over tokens of log(1+relu(MLM logits)))
    heads against the same computation done by hand on transformers' BERT states / masked-LM logits."""
    import transformers
    prompts = [[3, 17, 99, 250, 41, 7], list(range(10, 30))]
    hf = _hf_from_preset("bert-tiny", "BertForMaskedLM")
    sd = dict(hf.state_dict())
    g = torch.Generator().manual_seed(0)
    lin = torch.randn(24, 64, generator=g) * 0.2
    llm, m = _pooling_llm("bert-tiny", "HF_ColBERT", hf, colbert_dim=24)
    m.load_weights([("linear.weight", lin)])
    outs = llm.encode([{"prompt_token_ids": p} for p in prompts], pooling_task="token_embed")
    for p, o in zip(prompts, outs):
        with torch.no_grad():
            h = hf.bert(torch.tensor([p])).last_hidden_state[0]
        ref = torch.nn.functional.normalize(h @ lin.t(), dim=-1)
        torch.testing.assert_close(torch.as_tensor(o.outputs.data).float(), ref, atol=3e-4, rtol=1e-3)
    llm.shutdown()
    llm, m = _pooling_llm("bert-tiny", "BertSpladeSparseEmbeddingModel", hf)
    from vllm_b200 import PoolingParams
    outs = llm.encode([{"prompt_token_ids": p} for p in prompts], PoolingParams(task="embed", normalize=False))
    for p, o in zip(prompts, outs):
        with torch.no_grad():
            ref = torch.log1p(torch.relu(hf(torch.tensor([p])).logits[0])).amax(0)
        got = torch.as_tensor(o.outputs.data).float()
        assert got.shape == (512,)
        torch.testing.assert_close(got, ref, atol=3e-4, rtol=1e-3)
    llm.shutdown()


def test_bidirectional_llama_embeddings_match_hf():
    """A decoder checkpoint served as a bidirectional encoder (``LlamaBidirectionalModel``, mean pooling): equals
    transformers' LlamaModel run with an all-ones 4-D mask; the Voyage variant adds its linear head."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    kw = {k: v for k, v in cfg.items() if k not in ("architectures", "model_type", "torch_dtype")}
    hf_cfg = transformers.AutoConfig.for_model("llama", **kw)
    torch.manual_seed(1)
    base = transformers.LlamaModel(hf_cfg).float().eval()
    prompts = [[5, 17, 99, 3, 250, 41], list(range(10, 45))]

    def hf_states(p):
        n = len(p)
        with torch.no_grad():
            return base(torch.tensor([p]), attention_mask=torch.zeros(1, 1, n, n)).last_hidden_state[0]
    llm = make_llm("llama-tiny", hf_overrides={"architectures": ["LlamaBidirectionalModel"], "pooling": "avg"}, runner="pooling")
    assert llm.llm_engine.vllm_config.model_config.is_encoder_only
    llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model.load_weights(base.state_dict().items())
    for p, o in zip(prompts, llm.embed([{"prompt_token_ids": p} for p in prompts])):
        ref = torch.nn.functional.normalize(hf_states(p).mean(0), dim=-1)
        torch.testing.assert_close(torch.as_tensor(o.outputs.data).float(), ref, atol=2e-4, rtol=1e-3)
    llm.shutdown()


def test_bge_m3_three_views():
    """Dense / multi-vector / lexical outputs of BGE-M3 from one XLM-R encoder, against the heads applied by hand to
    transformers' hidden states."""
    hf = _hf_from_preset("bert-tiny", "XLMRobertaModel", "xlm-roberta", pad_token_id=1, max_position_embeddings=258)
    g = torch.Generator().manual_seed(0)
    heads = {"sparse_linear.weight": torch.randn(1, 64, generator=g) * 0.3, "sparse_linear.bias": torch.randn(1, generator=g) * 0.1,
             "colbert_linear.weight": torch.randn(64, 64, generator=g) * 0.2, "colbert_linear.bias": torch.randn(64, generator=g) * 0.1}
    llm, m = _pooling_llm("bert-tiny", "BgeM3EmbeddingModel", hf, model_type="xlm-roberta", max_position_embeddings=258)
    m.load_weights(heads.items())
    p = [3, 17, 99, 250, 41, 7, 12]
    with torch.no_grad():
        h = hf(torch.tensor([p]), position_ids=torch.arange(2, 2 + len(p))[None]).last_hidden_state[0]
    dense = torch.as_tensor(llm.embed([{"prompt_token_ids": p}])[0].outputs.data).float()
    torch.testing.assert_close(dense, torch.nn.functional.normalize(h[0], dim=-1), atol=3e-4, rtol=1e-3)
    multi = torch.as_tensor(llm.encode([{"prompt_token_ids": p}], pooling_task="token_embed")[0].outputs.data).float()
    ref = torch.nn.functional.normalize(h[1:] @ heads["colbert_linear.weight"].t() + heads["colbert_linear.bias"], dim=-1)
    torch.testing.assert_close(multi, ref, atol=3e-4, rtol=1e-3)
    lex = torch.as_tensor(llm.encode([{"prompt_token_ids": p}], pooling_task="token_classify")[0].outputs.data).float()
    ref = torch.relu(h @ heads["sparse_linear.weight"].t() + heads["sparse_linear.bias"]).squeeze(-1)
    torch.testing.assert_close(lex, ref, atol=3e-4, rtol=1e-3)
    llm.shutdown()


def test_nemotron_nas_blocks():
    """DeciLM / Nemotron-NAS: with uniform block configs the model IS the Llama it was searched from (same weights, same
    outputs); with a searched layout (different KV-head counts per layer, a layer without attention, a layer without MLP)
    it builds exactly the parameters the checkpoint has, serves from per-layer-sized KV caches, and generates."""
    from tests.utils import llama_full_weights
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=4))
    uniform = [{"attention": {"n_heads_in_group": 2, "no_op": False}, "ffn": {"intermediate_size": 256, "no_op": False}}] * 2
    _same(_gen("llama-tiny", w), _gen("llama-tiny", w, {"architectures": ["DeciLMForCausalLM"], "block_configs": uniform}))
    searched = [{"attention": {"n_heads_in_group": 1, "no_op": False}, "ffn": {"ffn_mult": 1.0, "no_op": False}},
                {"attention": {"no_op": True}, "ffn": {"intermediate_size": 128, "no_op": False}},
                {"attention": {"n_heads_in_group": 4, "no_op": False}, "ffn": {"no_op": True}},
                {"attention": {"no_op": True}, "ffn": {"no_op": True}}]
    llm = make_llm("llama-tiny", hf_overrides={"architectures": ["DeciLMForCausalLM"], "block_configs": searched, "num_hidden_layers": 4})
    m = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model
    names = {n for n, _ in m.named_parameters()}
    assert "model.layers.1.post_attention_layernorm.weight" in names and not any(n.startswith("model.layers.1.self_attn") for n in names)
    assert not any(n.startswith("model.layers.2.mlp") for n in names) and not any(n.startswith("model.layers.3.") for n in names)
    assert m.model.layers[0].self_attn.qkv_proj.num_kv_heads == 4 and m.model.layers[2].self_attn.qkv_proj.num_kv_heads == 1
    assert m.model.layers[0].mlp.gate_up_proj.weight.shape[0] == 2 * 256            # round_up(int(2 * 1.0 * 128 / 3), 256)
    out = llm.generate([{"prompt_token_ids": p} for p in PROMPTS], SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True))
    assert all(len(o.outputs[0].token_ids) == 6 for o in out)
    llm.shutdown()


@pytest.mark.parametrize("moe", [False, True])
def test_qwen3_vl_matches_hf(moe):
    """Qwen3-VL: resampled position table + 2-D rotary ViT, DeepStack features added after the first decoder layers,
    interleaved M-RoPE, q/k-normed decoder — greedy continuation equals transformers for image+text and text-only prompts."""
    import numpy as np
    import transformers
    from PIL import Image
    from vllm_b200.config import get_preset_hf_config
    preset = "qwen3vlmoe-tiny" if moe else "qwen3vl-tiny"
    cfg = get_preset_hf_config(preset)
    text_keys = ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "head_dim",
                 "num_key_value_heads", "max_position_embeddings", "rms_norm_eps", "hidden_act", "tie_word_embeddings", "attention_bias")
    if moe:
        text_keys += ("moe_intermediate_size", "num_experts", "num_experts_per_tok", "norm_topk_prob", "decoder_sparse_step", "mlp_only_layers")
    tc = {k: cfg[k] for k in text_keys}
    tc["rope_parameters"] = dict(cfg["rope_parameters"])
    CfgCls = transformers.Qwen3VLMoeConfig if moe else transformers.Qwen3VLConfig
    ModelCls = transformers.Qwen3VLMoeForConditionalGeneration if moe else transformers.Qwen3VLForConditionalGeneration
    hf_cfg = CfgCls(text_config=tc, vision_config=cfg["vision_config"], image_token_id=500, video_token_id=501,
                                        vision_start_token_id=502, vision_end_token_id=503, tie_word_embeddings=False)
    torch.manual_seed(0)
    hf = ModelCls(hf_cfg).float().eval()
    rng = np.random.default_rng(0)
    img = Image.fromarray(rng.integers(0, 255, (64, 96, 3), dtype=np.uint8))
    ip = transformers.Qwen2VLImageProcessor(patch_size=16, merge_size=2, image_mean=[0.5] * 3, image_std=[0.5] * 3,
                                            size={"shortest_edge": 32 * 32, "longest_edge": 16 * 16 * 4 * 4096})
    px = ip(images=[img], return_tensors="pt")
    n_img = int(px["image_grid_thw"][0].prod()) // 4
    pre, post = [5, 17, 99, 502], [503, 3, 250, 41, 7]
    ids_hf = pre + [500] * n_img + post
    with torch.no_grad():
        ref_img = hf.generate(input_ids=torch.tensor([ids_hf]), attention_mask=torch.ones(1, len(ids_hf), dtype=torch.long),
                              pixel_values=px["pixel_values"], image_grid_thw=px["image_grid_thw"],
                              mm_token_type_ids=torch.tensor([[1 if t == 500 else 0 for t in ids_hf]]), max_new_tokens=8,
                              do_sample=False, pad_token_id=0, eos_token_id=None)[0, len(ids_hf):].tolist()
    text_prompt = [5, 17, 99, 3, 250, 41, 5, 17, 99, 3]
    ref_txt = hf_greedy(hf, text_prompt, 8)
    llm = make_llm(preset, weights=hf.state_dict().items(), enable_prefix_caching=False, hf_overrides={"min_pixels": 32 * 32})
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    outs = llm.generate([{"prompt_token_ids": pre + [500] + post, "multi_modal_data": {"image": img}},
                         {"prompt_token_ids": text_prompt}], sp)
    llm.shutdown()
    assert len(outs[0].prompt_token_ids) == len(ids_hf)
    assert outs[0].outputs[0].token_ids == ref_img, (outs[0].outputs[0].token_ids, ref_img)
    assert outs[1].outputs[0].token_ids == ref_txt


def test_llava_onevision_matches_hf():
    """LLaVA-OneVision: SigLIP tiles, unpadded feature map, down-sampling once the map exceeds ``anyres_max_N`` tiles' worth
    of patches, newline column, Qwen2 decoder — for a 1×3 canvas (shrunk) and a 2×2 canvas, against transformers."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llavaov-tiny")
    drop = ("architectures", "model_type", "vision_config", "image_token_index", "vision_feature_layer", "vision_feature_select_strategy",
            "projector_hidden_act", "vision_aspect_ratio", "image_grid_pinpoints", "multimodal_projector_bias", "text_model_type", "torch_dtype")
    hf_cfg = transformers.LlavaOnevisionConfig(vision_config=cfg["vision_config"], text_config=dict(model_type="qwen2", **_vlm_text_cfg(cfg, drop)),
                                               image_token_index=511, vision_feature_layer=-1, vision_feature_select_strategy="full",
                                               vision_aspect_ratio="anyres_max_2", image_grid_pinpoints=cfg["image_grid_pinpoints"],
                                               multimodal_projector_bias=True, tie_word_embeddings=False)
    torch.manual_seed(0)
    hf = transformers.LlavaOnevisionForConditionalGeneration(hf_cfg).eval()
    with torch.no_grad():
        hf.model.image_newline.normal_(0, 0.5)
    llm = make_llm("llavaov-tiny", weights=hf.state_dict().items())
    g = torch.Generator().manual_seed(1)
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    for size, n_tiles in (((30, 95), 4), ((60, 64), 5)):
        tiles = torch.randn(n_tiles, 3, 32, 32, generator=g)
        with torch.no_grad():
            feats = hf.get_image_features(pixel_values=tiles[None], image_sizes=torch.tensor([size]), vision_feature_layer=-1,
                                          vision_feature_select_strategy="full").pooler_output
        n = feats[0].shape[0] if isinstance(feats, (list, tuple)) else feats.shape[0]
        exp_ids = [5, 17] + [511] * n + [99, 3, 250]
        with torch.no_grad():
            ref = hf.generate(input_ids=torch.tensor([exp_ids]), pixel_values=tiles[None], image_sizes=torch.tensor([size]),
                              max_new_tokens=8, do_sample=False, pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
        out = llm.generate([{"prompt_token_ids": [5, 17, 511, 99, 3, 250],
                             "multi_modal_data": {"image": {"tiles": tiles, "image_size": size}}}], sp)
        assert len(out[0].prompt_token_ids) == len(exp_ids), (size, n, len(out[0].prompt_token_ids) - 5)
        assert out[0].outputs[0].token_ids == ref, size
    llm.shutdown()


def test_paligemma_prefix_lm_matches_hf():
    """PaliGemma: image + prompt attend bidirectionally (prefix-LM), continuation is causal — greedy output equals
    transformers given ``token_type_ids`` that mark the whole prompt as prefix; the engine keeps the prompt in one pass."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("paligemma-tiny")
    drop = ("architectures", "model_type", "vision_config", "image_token_index", "text_model_type", "projection_dim", "torch_dtype", "hidden_act")
    hf_cfg = transformers.PaliGemmaConfig(vision_config=cfg["vision_config"], text_config=dict(model_type="gemma", **_vlm_text_cfg(cfg, drop)),
                                          image_token_index=511, projection_dim=64, hidden_size=64)
    torch.manual_seed(0)
    hf = transformers.PaliGemmaForConditionalGeneration(hf_cfg).eval()
    g = torch.Generator().manual_seed(1)
    img = torch.randn(3, 32, 32, generator=g)
    exp_ids = [511] * 4 + [2, 17, 99, 3, 250]                      # image tokens first, then <bos> + prompt
    with torch.no_grad():
        ref = hf.generate(input_ids=torch.tensor([exp_ids]), pixel_values=img[None], token_type_ids=torch.zeros(1, len(exp_ids), dtype=torch.long),
                          attention_mask=torch.ones(1, len(exp_ids), dtype=torch.long), max_new_tokens=8, do_sample=False,
                          pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
    llm = make_llm("paligemma-tiny", weights=hf.state_dict().items())
    vc = llm.llm_engine.vllm_config
    assert not vc.scheduler_config.enable_chunked_prefill and not vc.cache_config.enable_prefix_caching
    out = llm.generate([{"prompt_token_ids": [511, 2, 17, 99, 3, 250], "multi_modal_data": {"image": img}},
                        {"prompt_token_ids": [511, 2, 9, 9, 41], "multi_modal_data": {"image": img}}],
                       SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True))
    assert out[0].outputs[0].token_ids == ref
    llm.shutdown()


def test_hyperclovax_is_granite_with_optional_output_norms():
    """HyperCLOVA X without post-norms equals Granite (logits_scaling multiplies instead of divides); with ``use_post_norm`` it
    is the sandwich block: identity-initialised output norms leave unit-norm'd sub-layer outputs, so the two differ."""
    from tests.utils import llama_full_weights
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=6))
    hd = cfg["hidden_size"] // cfg["num_attention_heads"]
    common = {"embedding_multiplier": 4.0, "residual_multiplier": 0.5, "attention_multiplier": hd ** -0.5}
    a = _gen("llama-tiny", w, dict(common, architectures=["GraniteForCausalLM"], logits_scaling=4.0))
    b = _gen("llama-tiny", w, dict(common, architectures=["HyperCLOVAXForCausalLM"], logits_scaling=0.25))
    _same(a, b)
    c = _gen("llama-tiny", w, dict(common, architectures=["HyperCLOVAXForCausalLM"], logits_scaling=0.25, use_post_norm=True))
    assert [t for t, _ in c] != [t for t, _ in b]
    from vllm_b200.models.registry import resolve_model_cls
    assert resolve_model_cls("Rnj1ForCausalLM").__name__ == "Gemma3ForCausalLM"


def test_ultravox_projector_and_splice():
    """Ultravox (no transformers implementation): the tower is the Whisper encoder already checked against transformers, so
    the test pins the projector (frame stacking with tail padding, RMSNorm → linear → SwiGLU → RMSNorm → linear) and the
    token count against a hand computation, and runs the spliced model end to end."""
    llm = make_llm("ultravox-tiny")
    m = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for p in m.multi_modal_projector.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.2 if p.dim() > 1 else 1.0) + (0.0 if p.dim() > 1 else 1.0))
    x = torch.randn(10, 32, generator=g)                            # 10 encoder frames, stack 4 → 3 tokens (2 padded rows)
    P = m.multi_modal_projector

    def rms(v, w):
        return v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + 1e-6) * w
    st = torch.cat([x, torch.zeros(2, 32)]).view(3, 128)
    h = rms(st, P["ln_pre"].weight) @ P["linear_1"].weight.t()
    a, b = h.chunk(2, dim=-1)
    ref = rms(a * torch.nn.functional.silu(b), P["ln_mid"].weight) @ P["linear_2"].weight.t()
    torch.testing.assert_close(m._project(x), ref, atol=1e-5, rtol=1e-4)
    assert type(m).num_tokens({}, 37, 4) == 5 and type(m).num_tokens({}, 1, 4) == 1      # ceil(((37-1)//2+1) / 4)
    mel = torch.randn(16, 37, generator=g)
    out = llm.generate([{"prompt_token_ids": [5, 17, 511, 99], "multi_modal_data": {"audio": mel}}],
                       SamplingParams(max_tokens=4, temperature=0.0, ignore_eos=True))
    assert len(out[0].prompt_token_ids) == 3 + 5 and len(out[0].outputs[0].token_ids) == 4
    llm.shutdown()


def test_glm_moe_dsa_dense_path_matches_hf_within_index_topk():
    """GLM-5 / DeepSeek-V3.2 (sparse-attention indexer): while the context fits in ``index_topk`` the indexer selects every
    token, so our dense MLA path must equal transformers (eager attention: its sdpa path drops the causal mask when the
    mask argument is absent).  Indexer tensors in the checkpoint are skipped; V3.2 keeps DeepSeek's interleaved rotary."""
    hf = make_hf_model("glmmoedsa-tiny")
    hf.config._attn_implementation = "eager"
    for layer in hf.model.layers:
        layer.self_attn.config._attn_implementation = "eager"
    sd = hf.state_dict()
    assert any(".indexer." in k for k in sd)
    llm = make_llm("glmmoedsa-tiny", weights=sd.items())
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    for p, o in zip(PROMPTS, llm.generate([{"prompt_token_ids": p} for p in PROMPTS], sp)):
        assert o.outputs[0].token_ids == hf_greedy(hf, p, 8)
    m = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model
    assert m.model.layers[0].self_attn.rope.is_neox_style
    llm.shutdown()
    llm = make_llm("glmmoedsa-tiny", hf_overrides={"architectures": ["DeepseekV32ForCausalLM"]})
    m = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model
    assert not m.model.layers[0].self_attn.rope.is_neox_style
    llm.shutdown()


def test_transformers_backend_pooling_and_modernbert_colbert():
    """``TransformersEmbeddingModel`` (the pooling adapter over the Transformers back-end) gives the same last-token
    embedding as the native ``LlamaModel`` adapter on the same weights; ColBERT-ModernBERT applies its projection per token."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    kw = {k: v for k, v in cfg.items() if k not in ("architectures", "model_type", "torch_dtype")}
    torch.manual_seed(1)
    base = transformers.LlamaModel(transformers.AutoConfig.for_model("llama", **kw)).float().eval()
    prompts = [[5, 17, 99, 3, 250, 41], list(range(10, 45))]
    outs = {}
    for arch in ("LlamaModel", "TransformersEmbeddingModel"):
        llm = make_llm("llama-tiny", hf_overrides={"architectures": [arch]}, runner="pooling")
        llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model.load_weights(base.state_dict().items())
        outs[arch] = [torch.as_tensor(o.outputs.data).float() for o in llm.embed([{"prompt_token_ids": p} for p in prompts])]
        llm.shutdown()
    for a, b in zip(outs["LlamaModel"], outs["TransformersEmbeddingModel"]):
        torch.testing.assert_close(a, b, atol=3e-4, rtol=1e-3)
    hf = make_hf_model("modernbert-tiny")
    g = torch.Generator().manual_seed(0)
    lin = torch.randn(16, 64, generator=g) * 0.2
    llm, m = _pooling_llm("modernbert-tiny", "ColBERTModernBertModel", hf, colbert_dim=16)
    m.load_weights([("linear.weight", lin)])
    p = [3, 17, 99, 250, 41, 7]
    got = torch.as_tensor(llm.encode([{"prompt_token_ids": p}], pooling_task="token_embed")[0].outputs.data).float()
    with torch.no_grad():
        ref = torch.nn.functional.normalize(hf(torch.tensor([p])).last_hidden_state[0] @ lin.t(), dim=-1)
    torch.testing.assert_close(got, ref, atol=3e-4, rtol=1e-3)
    llm.shutdown()


def test_qwen3_5_vl_matches_hf():
    """Qwen3.5-VL: Qwen3-VL tower (no DeepStack) + gated-delta-net hybrid decoder with interleaved, partial M-RoPE on its
    attention layers — greedy continuation for an image prompt equals transformers."""
    import numpy as np
    import transformers
    from PIL import Image
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("qwen35vl-tiny")
    drop = ("architectures", "model_type", "vision_config", "image_token_id", "video_token_id", "vision_start_token_id",
            "vision_end_token_id", "torch_dtype")
    tc = {k: v for k, v in cfg.items() if k not in drop}
    hf_cfg = transformers.Qwen3_5Config(text_config=tc, vision_config=cfg["vision_config"], image_token_id=500, video_token_id=501,
                                        vision_start_token_id=502, vision_end_token_id=503, tie_word_embeddings=False)
    torch.manual_seed(0)
    hf = transformers.Qwen3_5ForConditionalGeneration(hf_cfg).float().eval()
    rng = np.random.default_rng(0)
    img = Image.fromarray(rng.integers(0, 255, (64, 96, 3), dtype=np.uint8))
    ip = transformers.Qwen2VLImageProcessor(patch_size=16, merge_size=2, image_mean=[0.5] * 3, image_std=[0.5] * 3,
                                            size={"shortest_edge": 32 * 32, "longest_edge": 16 * 16 * 4 * 4096})
    px = ip(images=[img], return_tensors="pt")
    n_img = int(px["image_grid_thw"][0].prod()) // 4
    pre, post = [5, 17, 99, 502], [503, 3, 250, 41, 7]
    ids_hf = pre + [500] * n_img + post
    with torch.no_grad():
        ref = hf.generate(input_ids=torch.tensor([ids_hf]), attention_mask=torch.ones(1, len(ids_hf), dtype=torch.long),
                          pixel_values=px["pixel_values"], image_grid_thw=px["image_grid_thw"],
                          mm_token_type_ids=torch.tensor([[1 if t == 500 else 0 for t in ids_hf]]), max_new_tokens=8,
                          do_sample=False, pad_token_id=0, eos_token_id=None)[0, len(ids_hf):].tolist()
    llm = make_llm("qwen35vl-tiny", weights=hf.state_dict().items(), enable_prefix_caching=False, hf_overrides={"min_pixels": 32 * 32})
    out = llm.generate([{"prompt_token_ids": pre + [500] + post, "multi_modal_data": {"image": img}}],
                       SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True))
    llm.shutdown()
    assert out[0].outputs[0].token_ids == ref, (out[0].outputs[0].token_ids, ref)


def test_llava_next_video_matches_hf():
    """LLaVA-NeXT-Video: per-frame tower features, stride-2 average pooling, projector, frames concatenated over the video
    placeholder — greedy continuation equals transformers for a 3-frame clip."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = dict(get_preset_hf_config("llava-tiny"))
    pins = [[28, 28], [28, 56], [56, 28]]
    drop = ("architectures", "model_type", "vision_config", "image_token_index", "vision_feature_layer",
            "vision_feature_select_strategy", "projector_hidden_act", "torch_dtype")
    hf_cfg = transformers.LlavaNextVideoConfig(vision_config=cfg["vision_config"], text_config=dict(model_type="llama", **_vlm_text_cfg(cfg, drop)),
                                               image_token_index=511, video_token_index=510, vision_feature_layer=-2,
                                               vision_feature_select_strategy="default", projector_hidden_act="gelu",
                                               image_grid_pinpoints=pins, spatial_pool_mode="average", spatial_pool_stride=2)
    torch.manual_seed(0)
    hf = transformers.LlavaNextVideoForConditionalGeneration(hf_cfg).eval()
    g = torch.Generator().manual_seed(1)
    frames = torch.randn(3, 3, 28, 28, generator=g)
    n = 3 * 1                                              # (28 / 14 / 2)^2 = 1 token per frame
    exp_ids = [5, 17] + [510] * n + [99, 3, 250]
    with torch.no_grad():
        ref = hf.generate(input_ids=torch.tensor([exp_ids]), pixel_values_videos=frames[None], max_new_tokens=8, do_sample=False,
                          pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
    llm = make_llm("llava-tiny", weights=hf.state_dict().items(),
                   hf_overrides={"architectures": ["LlavaNextVideoForConditionalGeneration"], "image_grid_pinpoints": pins,
                                 "video_token_index": 510, "spatial_pool_stride": 2, "spatial_pool_mode": "average"})
    out = llm.generate([{"prompt_token_ids": [5, 17, 510, 99, 3, 250], "multi_modal_data": {"video": frames}}],
                       SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True))
    assert out[0].prompt_token_ids == exp_ids
    assert out[0].outputs[0].token_ids == ref
    llm.shutdown()


def test_minicpm3_is_dense_mla_with_mup_scalars():
    """MiniCPM3 (no transformers implementation): with neutral scalars (scale_emb 1, scale_depth √L, dim_model_base = hidden)
    it must equal the dense DeepSeek MLA decoder with rotate-half rotary on the same weights; non-neutral scalars change
    the output."""
    import math
    base = {"n_routed_experts": None, "first_k_dense_replace": 99, "rope_interleave": False}
    a = _gen("deepseek-tiny", None, dict(base, architectures=["DeepseekV3ForCausalLM"]))
    neutral = {"architectures": ["MiniCPM3ForCausalLM"], "n_routed_experts": None, "scale_emb": 1.0, "scale_depth": math.sqrt(3),
               "dim_model_base": 128}
    b = _gen("deepseek-tiny", None, neutral)
    _same(a, b)
    c = _gen("deepseek-tiny", None, dict(neutral, scale_emb=12.0, scale_depth=1.4, dim_model_base=32))
    # (tiny dummy weights keep the argmax; the scalars must show in the log-probabilities)
    assert max(abs(x - y) for (_, lc), (_, lb) in zip(c, b) for x, y in zip(lc, lb)) > 1e-4


def test_llama4_vision_matches_hf():
    """Llama-4 multimodal: unfolded-patch ViT with 2-D rotary pairs, pixel-shuffle adapter, projector, Llama-4 MoE decoder
    (chunked attention, NoPE layers) — greedy continuation with an image equals transformers."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama4mm-tiny")
    drop = ("architectures", "model_type", "vision_config", "image_token_index", "torch_dtype")
    hf_cfg = transformers.Llama4Config(vision_config=cfg["vision_config"], text_config=_vlm_text_cfg(cfg, drop), image_token_index=511,
                                       boi_token_index=509, eoi_token_index=510)
    torch.manual_seed(0)
    hf = transformers.Llama4ForConditionalGeneration(hf_cfg).float().eval()
    g = torch.Generator().manual_seed(1)
    img = torch.randn(3, 32, 32, generator=g)
    exp_ids = [5, 17] + [511] * 4 + [99, 3, 250]
    with torch.no_grad():
        ref = hf.generate(input_ids=torch.tensor([exp_ids]), pixel_values=img[None], max_new_tokens=8, do_sample=False,
                          pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
    llm = make_llm("llama4mm-tiny", weights=hf.state_dict().items())
    out = llm.generate([{"prompt_token_ids": [5, 17, 511, 99, 3, 250], "multi_modal_data": {"image": img}}],
                       SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True))
    assert out[0].prompt_token_ids == exp_ids
    assert out[0].outputs[0].token_ids == ref
    llm.shutdown()


def test_llava_with_pixtral_tower_matches_hf():
    """pixtral-12b in the HF layout = LLaVA with a Pixtral tower: 2-D rotary ViT, all patch tokens, row-break image layout —
    greedy continuation equals transformers' LlavaForConditionalGeneration."""
    import transformers
    from vllm_b200.config import get_preset_hf_config
    cfg = dict(get_preset_hf_config("llava-tiny"))
    vc = dict(model_type="pixtral", hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2, head_dim=16,
              image_size=32, patch_size=8, hidden_act="silu", rope_theta=10000.0)
    drop = ("architectures", "model_type", "vision_config", "image_token_index", "vision_feature_layer",
            "vision_feature_select_strategy", "projector_hidden_act", "torch_dtype")
    hf_cfg = transformers.LlavaConfig(vision_config=vc, text_config=dict(model_type="llama", **_vlm_text_cfg(cfg, drop)), image_token_index=10,
                                      vision_feature_layer=-1, vision_feature_select_strategy="full", projector_hidden_act="gelu")
    torch.manual_seed(0)
    hf = transformers.LlavaForConditionalGeneration(hf_cfg).eval()
    g = torch.Generator().manual_seed(1)
    img = torch.randn(3, 32, 32, generator=g)
    rows = []
    for r in range(4):
        rows += [10] * 4 + [13 if r == 3 else 12]
    exp_ids = [5, 17] + rows + [99, 3, 250]
    with torch.no_grad():
        ref = hf.generate(input_ids=torch.tensor([exp_ids]), pixel_values=img[None], image_sizes=torch.tensor([[32, 32]]),
                          max_new_tokens=8, do_sample=False, pad_token_id=0, eos_token_id=None)[0, len(exp_ids):].tolist()
    llm = make_llm("llava-tiny", weights=hf.state_dict().items(),
                   hf_overrides={"vision_config": vc, "image_token_index": 10, "vision_feature_layer": -1,
                                 "vision_feature_select_strategy": "full"})
    out = llm.generate([{"prompt_token_ids": [5, 17, 10, 99, 3, 250], "multi_modal_data": {"image": img}}],
                       SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True))
    assert out[0].prompt_token_ids == exp_ids
    assert out[0].outputs[0].token_ids == ref
    llm.shutdown()


def test_solar_block_skip_connections():
    """Solar (depth-up-scaled Llama with block skip connections): blend weight 0 is the plain Llama; a non-zero weight must
    equal a by-hand blend — checked through the invariant that blending a layer's input with ITSELF (save and mix at the
    same layer) is the identity, and that mixing an earlier stream changes the output."""
    from tests.utils import llama_full_weights
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=8))
    ref = _gen("llama-tiny", w)
    solar = {"architectures": ["SolarForCausalLM"], "bskcn_1": [0], "bskcn_2": [], "bskcn_3": [1], "bskcn_4": []}
    _same(ref, _gen("llama-tiny", w, dict(solar, bskcn_tv=[0.9, 0.0])))                 # inference weight 0 → plain Llama
    _same(ref, _gen("llama-tiny", w, dict(solar, bskcn_3=[0], bskcn_tv=[0.9, 0.4])))     # stream blended with itself
    mixed = _gen("llama-tiny", w, dict(solar, bskcn_tv=[0.9, 0.4]))
    assert max(abs(x - y) for (_, a), (_, b) in zip(mixed, ref) for x, y in zip(a, b)) > 1e-4


def test_model_impl_selects_backend():
    """--model-impl: ``transformers`` serves a natively supported architecture through the HF back-end (same tokens), the
    native-only setting refuses an architecture that only the back-end could run."""
    from tests.utils import llama_full_weights
    from vllm_b200.config import get_preset_hf_config
    from vllm_b200.models.registry import resolve_model_cls
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=9))
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    outs = {}
    for impl in ("auto", "transformers"):
        llm = make_llm("llama-tiny", weights=w, model_impl=impl)
        name = type(llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model).__name__
        assert name == ("TransformersForCausalLM" if impl == "transformers" else "LlamaForCausalLM")
        outs[impl] = [o.outputs[0].token_ids for o in llm.generate([{"prompt_token_ids": p} for p in PROMPTS], sp)]
        llm.shutdown()
    assert outs["auto"] == outs["transformers"]
    with pytest.raises(ValueError, match="no native implementation"):
        resolve_model_cls("PersimmonForCausalLM", "vllm_b200")
    assert resolve_model_cls("PersimmonForCausalLM").__name__ == "TransformersForCausalLM"
