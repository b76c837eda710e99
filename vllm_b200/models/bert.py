"""Encoder-only models for embeddings / classification: BERT, RoBERTa.

Reference parity: vllm/model_executor/models/bert.py (``BertEmbedding``, ``BertLayer`` post-LN blocks,
``BertModel``, ``BertEmbeddingModel``, ``BertForSequenceClassification``) and roberta.py (position offset).
Bidirectional attention runs through the same paged-attention op with ``causal=False``; the whole sequence is
scheduled in one chunk (the engine disables chunked prefill for pooling-only models).
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import Activation, LayerNorm, VocabParallelEmbedding
from vllm_b200.layers.linear import ColumnParallelLinear, QKVParallelLinear, RowParallelLinear


class BertEmbedding(nn.Module):
    def __init__(self, cfg: dict, dtype, pos_offset: int = 0) -> None:
        super().__init__()
        H = cfg["hidden_size"]
        self.word_embeddings = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.position_embeddings = nn.Embedding(cfg["max_position_embeddings"], H, dtype=dtype)
        self.token_type_embeddings = nn.Embedding(cfg.get("type_vocab_size", 2), H, dtype=dtype)
        self.LayerNorm = LayerNorm(H, cfg.get("layer_norm_eps", 1e-12), dtype=dtype)
        self.pos_offset = pos_offset

    def forward(self, ids: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        x = (self.word_embeddings(ids) + self.position_embeddings(positions + self.pos_offset)
             + self.token_type_embeddings(torch.zeros_like(ids)))
        return self.LayerNorm(x)


class BertLayer(nn.Module):
    def __init__(self, cfg: dict, dtype, quant_config, prefix: str) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        hd = H // nh
        eps = cfg.get("layer_norm_eps", 1e-12)
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nh, True, dtype, quant_config, f"{prefix}.attention.self.qkv_proj")
        self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads,
                              prefix=f"{prefix}.attention.self.attn", causal=False)
        self.attn_out = RowParallelLinear(H, H, True, dtype=dtype, quant_config=quant_config,
                                          prefix=f"{prefix}.attention.output.dense")
        self.attn_ln = LayerNorm(H, eps, dtype=dtype)
        self.inter = ColumnParallelLinear(H, cfg["intermediate_size"], True, dtype=dtype, quant_config=quant_config,
                                          prefix=f"{prefix}.intermediate.dense")
        self.act = Activation(cfg.get("hidden_act", "gelu"))
        self.out = RowParallelLinear(cfg["intermediate_size"], H, True, dtype=dtype, quant_config=quant_config,
                                     prefix=f"{prefix}.output.dense")
        self.out_ln = LayerNorm(H, eps, dtype=dtype)
        self.q_size = self.qkv_proj.q_size

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        q, k, v = self.qkv_proj(x).split([self.q_size] * 3, dim=-1)
        x = self.attn_ln(self.attn_out(self.attn(q, k, v)) + x)
        return self.out_ln(self.out(self.act(self.inter(x))) + x)


class BertModel(nn.Module):
    is_pooling_model = True
    pos_offset = 0

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        self.embeddings = BertEmbedding(cfg, dtype, self.pos_offset)
        self.layers = nn.ModuleList([BertLayer(cfg, dtype, None, f"{prefix}encoder.layer.{i}")
                                     for i in range(cfg["num_hidden_layers"])])
        n_labels = cfg.get("num_labels") or len(cfg.get("id2label") or {}) or 0
        self.classifier = nn.Linear(cfg["hidden_size"], n_labels, dtype=dtype) if cfg.get("_classifier") else None
        self.pooler_dense = nn.Linear(cfg["hidden_size"], cfg["hidden_size"], dtype=dtype) if cfg.get("_classifier") else None

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = self.embeddings(input_ids, positions)
        for layer in self.layers:
            x = layer(x)
        return x

    def embed_input_ids(self, ids):
        return self.embeddings.word_embeddings(ids)

    def score(self, pooled: torch.Tensor) -> torch.Tensor:
        """Classification head on the pooled ([CLS]) state: tanh pooler + linear classifier."""
        if self.classifier is None:
            raise RuntimeError("this checkpoint has no classification head")
        return self.classifier(torch.tanh(self.pooler_dense(pooled)))

    def compute_logits(self, hidden):  # pragma: no cover - pooling-only
        raise RuntimeError("encoder-only model: use LLM.embed / LLM.classify")

    # HF name → ours
    _MAP = [("attention.self.query", "qkv_proj", "q"), ("attention.self.key", "qkv_proj", "k"),
            ("attention.self.value", "qkv_proj", "v")]
    _RENAME = [("attention.output.dense", "attn_out"), ("attention.output.LayerNorm", "attn_ln"),
               ("intermediate.dense", "inter"), ("output.dense", "out"), ("output.LayerNorm", "out_ln")]

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()
        for name, w in weights:
            for pre in ("bert.", "roberta.", "model."):
                if name.startswith(pre):
                    name = name[len(pre):]
            name = name.replace("encoder.layer.", "layers.").replace("gamma", "weight").replace("beta", "bias")
            if name.startswith("pooler.dense."):
                name = name.replace("pooler.dense.", "pooler_dense.")
            shard = None
            for hf, ours, sid in self._MAP:
                if f".{hf}." in name:
                    name, shard = name.replace(hf, ours), sid
                    break
            else:
                for hf, ours in self._RENAME:
                    if f".{hf}." in name:
                        name = name.replace(hf, ours)
                        break
            if name not in params:
                continue
            mod = modules[name.rsplit(".", 1)[0] if "." in name else ""]
            loader = getattr(mod, "weight_loader", None)
            if loader is not None:
                loader(params[name], w, shard)
            else:
                params[name].data.copy_(w)
            loaded.add(name)
        return loaded


class BertEmbeddingModel(BertModel):
    default_pooling = "CLS"


class RobertaModel(BertModel):
    pos_offset = 2          # RoBERTa position ids start at padding_idx + 1
    default_pooling = "CLS"


class BertForSequenceClassification(BertModel):
    default_pooling = "CLS"

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        vllm_config.model_config.hf_config["_classifier"] = True
        super().__init__(vllm_config, prefix)


class RobertaForSequenceClassification(RobertaModel):
    """RoBERTa / XLM-R classification head: dense → tanh → out_proj on the ``<s>`` state (no BERT pooler)
    (reference: models/roberta.py ``RobertaClassificationHead``)."""
    default_pooling = "CLS"

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__(vllm_config, prefix)
        cfg, dtype = self.cfg, vllm_config.model_config.torch_dtype
        n = int(cfg.get("num_labels") or len(cfg.get("id2label") or {}) or 1)
        self.cls_dense = nn.Linear(cfg["hidden_size"], cfg["hidden_size"], dtype=dtype)
        self.cls_out = nn.Linear(cfg["hidden_size"], n, dtype=dtype)

    def score(self, pooled: torch.Tensor) -> torch.Tensor:
        return self.cls_out(torch.tanh(self.cls_dense(pooled)))

    def load_weights(self, weights):
        return super().load_weights((n.replace("classifier.dense.", "cls_dense.").replace("classifier.out_proj.", "cls_out."), w)
                                    for n, w in weights)


class BertForTokenClassification(BertModel):
    """One label distribution per token: linear classifier on every position (``LLM.encode(..., pooling_task=
    "token_classify")``) (reference: models/bert.py ``BertForTokenClassification``)."""
    default_pooling = "ALL"

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__(vllm_config, prefix)
        cfg = self.cfg
        n = int(cfg.get("num_labels") or len(cfg.get("id2label") or {}) or 2)
        self.tok_classifier = nn.Linear(cfg["hidden_size"], n, dtype=vllm_config.model_config.torch_dtype)

    def score(self, hidden: torch.Tensor) -> torch.Tensor:
        return self.tok_classifier(hidden)

    def load_weights(self, weights):
        return super().load_weights((("tok_classifier." + n[len("classifier."):]) if n.startswith("classifier.") else n, w)
                                    for n, w in weights)


class RobertaForTokenClassification(BertForTokenClassification):
    pos_offset = 2


class _LateInteraction:
    """ColBERT-style late interaction: every token state goes through a bias-free linear map to a small dimension and is
    L2-normalised; queries and documents are scored by MaxSim outside the engine (``LLM.encode(..., pooling_task=
    "token_embed")``) (reference: models/colbert.py / bert.py ``HF_ColBERT``)."""

    def _init_colbert(self, cfg, dtype) -> None:
        dim = int(cfg.get("colbert_dim") or cfg.get("dim") or 128)
        self.colbert_linear = nn.Linear(cfg["hidden_size"], dim, bias=False, dtype=dtype)

    def project_embedding(self, x: torch.Tensor, per_token: bool) -> torch.Tensor:
        return self.colbert_linear(x) if per_token else x

    @staticmethod
    def _colbert_names(weights):
        for n, w in weights:
            if n in ("linear.weight", "colbert_linear.weight"):
                yield "colbert_linear.weight", w
            else:
                yield n, w


class HF_ColBERT(_LateInteraction, BertModel):
    default_pooling = "ALL"

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        BertModel.__init__(self, vllm_config, prefix)
        self._init_colbert(self.cfg, vllm_config.model_config.torch_dtype)

    def load_weights(self, weights):
        return BertModel.load_weights(self, self._colbert_names(weights))


class ColBERTJinaRobertaModel(HF_ColBERT):
    pos_offset = 2


class BertSpladeSparseEmbeddingModel(BertModel):
    """SPLADE: masked-LM logits per token → log(1 + relu) → max over the sequence = one sparse |V|-dimensional lexical
    vector (reference: models/bert.py ``BertSpladeSparseEmbeddingModel``)."""
    default_pooling = "ALL"

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__(vllm_config, prefix)
        cfg, dtype = self.cfg, vllm_config.model_config.torch_dtype
        H = cfg["hidden_size"]
        self.mlm_dense = nn.Linear(H, H, dtype=dtype)
        self.mlm_ln = nn.LayerNorm(H, cfg.get("layer_norm_eps", 1e-12), dtype=dtype)
        self.mlm_bias = nn.Parameter(torch.zeros(cfg["vocab_size"], dtype=dtype), requires_grad=False)
        self.mlm_decoder = None            # tied to the word embeddings unless the checkpoint carries its own matrix

    def project_embedding(self, x: torch.Tensor, per_token: bool) -> torch.Tensor:
        h = self.mlm_ln(torch.nn.functional.gelu(self.mlm_dense(x)))
        w = self.mlm_decoder if self.mlm_decoder is not None else self.embeddings.word_embeddings.weight
        logits = h @ w.t() + self.mlm_bias
        return torch.log1p(torch.relu(logits)).amax(dim=0) if per_token else torch.log1p(torch.relu(logits))

    def load_weights(self, weights):
        def ren():
            for n, w in weights:
                n = (n.replace("cls.predictions.transform.dense.", "mlm_dense.").replace("cls.predictions.transform.LayerNorm.", "mlm_ln.")
                     .replace("cls.predictions.bias", "mlm_bias").replace("cls.predictions.decoder.bias", "mlm_bias"))
                if n == "cls.predictions.decoder.weight":
                    if w.data_ptr() != 0 and not torch.equal(w, self.embeddings.word_embeddings.weight.to(w.dtype)):
                        self.mlm_decoder = nn.Parameter(w.clone().to(self.mlm_bias.dtype), requires_grad=False)
                    continue
                yield n, w
        return super().load_weights(ren())


class BgeM3EmbeddingModel(RobertaModel):
    """BGE-M3: one XLM-R encoder, three retrieval views — dense ([CLS], ``embed``), multi-vector (``colbert_linear`` on every
    token after [CLS], ``token_embed``) and lexical (``relu(sparse_linear)`` weight per token, ``token_classify``)
    (reference: models/roberta.py ``BgeM3EmbeddingModel``)."""
    default_pooling = "CLS"
    raw_scores = True

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__(vllm_config, prefix)
        cfg, dtype = self.cfg, vllm_config.model_config.torch_dtype
        H = cfg["hidden_size"]
        self.sparse_linear = nn.Linear(H, 1, dtype=dtype)
        self.colbert_linear = nn.Linear(H, int(cfg.get("colbert_dim") or H), dtype=dtype)

    def project_embedding(self, x: torch.Tensor, per_token: bool) -> torch.Tensor:
        return self.colbert_linear(x[1:]) if per_token else x

    def score(self, hidden: torch.Tensor) -> torch.Tensor:
        return torch.relu(self.sparse_linear(hidden)).squeeze(-1)
