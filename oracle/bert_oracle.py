"""CPU oracle for the BERT encoder hot path (TEST INFRASTRUCTURE ONLY).

This file is the *checker*, not the product.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.  The product path (`spokennlp_b200/`) never does.

It restates, in explicit fp32 tensor arithmetic on the CPU (no `nn.Module`s, no
HuggingFace code), the algorithm the reference delegates to
`transformers.models.bert.modeling_bert.BertModel`, following the only in-tree
written statement of that arithmetic:

  * embeddings ............ mmvts/src/models/cross_encoder/bert_model.py:166-210
  * self-attention ........ mmvts/src/models/cross_encoder/bert_model.py:213-361
      (scores / sqrt(d) *then* + additive mask, softmax over keys, P @ V)
  * attention output ...... mmvts/src/models/cross_encoder/bert_model.py:364-375
  * intermediate (GELU) ... mmvts/src/models/cross_encoder/bert_model.py:427-439
      (`ACT2FN["gelu"]` is the exact erf GELU)
  * output ................ mmvts/src/models/cross_encoder/bert_model.py:442-453
  * one layer ............. mmvts/src/models/cross_encoder/bert_model.py:518-553
  * cross layer ........... mmvts/src/models/cross_encoder/bert_model.py:456-515
  * pooler ................ mmvts/src/models/cross_encoder/bert_model.py:689-701
  * token-cls head + CE ... emnlp2023-topic_segmentation/src/models/modules/loss_calculator.py:17,42-44
                            emnlp2023-topic_segmentation/src/models/modules/utils.py:173-182
  * ditto pooling ......... ditto/evaluation_ditto.py:125-155

Parity pin: `oracle/make_goldens.py` runs the real reference (HF `BertModel`
eager + the reference's own wrappers imported from /root/reference) in the build
container and commits its outputs under `tests/golden/`; `tests/test_oracle.py`
checks this restatement against those vectors.  The reference itself holds no
golden vectors for this path, so the pin is "outputs of the
reference run here", not "reference test fixtures".

All functions take a flat `state_dict`-style mapping with HuggingFace key names
so reference checkpoints drive the oracle unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor


@dataclass
class OracleConfig:
    hidden_size: int = 768
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    num_hidden_layers: int = 12
    layer_norm_eps: float = 1e-12
    vocab_size: int = 30522
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    pad_token_id: Optional[int] = 0      # nn.Embedding(padding_idx=config.pad_token_id): bert_model.py:171 (HF BertConfig default 0)

    @classmethod
    def from_hf(cls, cfg) -> "OracleConfig":
        return cls(**{f: getattr(cfg, f) for f in cls.__dataclass_fields__ if hasattr(cfg, f)})


# ----------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------

def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """y = x @ w.T + b   (torch.nn.Linear convention: w is [out, in])."""
    y = x @ w.t()
    return y if b is None else y + b


def layer_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float) -> Tensor:
    """Biased-variance LayerNorm over the last dim (bert_model.py:180,368,446)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * gamma + beta


def gelu_erf(x: Tensor) -> Tensor:
    """Exact GELU: 0.5 x (1 + erf(x / sqrt(2)))  (HF ACT2FN['gelu'])."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def additive_key_mask(attention_mask: Optional[Tensor], dtype: torch.dtype,
                      fill: Optional[float] = None) -> Optional[Tensor]:
    """[B,S] 0/1 mask -> additive [B,1,1,S] mask.

    HF 5.x fills with finfo.min; mmvts uses -1e6 (ma_encoder.py:55-56); the TF
    fork uses -1e4.  After the softmax's max-subtraction they all give exactly
    0 probability to masked keys whenever a row has one valid key, which is the
    only case the reference produces.
    """
    if attention_mask is None:
        return None
    if fill is None:
        fill = torch.finfo(dtype).min
    m = attention_mask.to(dtype)
    return ((1.0 - m) * fill)[:, None, None, :]


# ----------------------------------------------------------------------------
# model pieces
# ----------------------------------------------------------------------------

def embeddings(sd: Dict[str, Tensor], cfg: OracleConfig, input_ids: Tensor,
               token_type_ids: Optional[Tensor] = None,
               position_ids: Optional[Tensor] = None,
               inputs_embeds: Optional[Tensor] = None,
               prefix: str = "embeddings.", masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """LN(word[ids] + type[tt] + pos[pos]) — bert_model.py:184-210; `masks["emb"]` = the dropout multiplier of :209.
    The word table is an nn.Embedding with padding_idx = pad_token_id (:171): the pad row is read like any other in the
    forward and receives NO gradient in the backward (F.embedding's padding_idx does exactly that)."""
    if inputs_embeds is None:
        inputs_embeds = torch.nn.functional.embedding(input_ids, sd[prefix + "word_embeddings.weight"],
                                                      padding_idx=getattr(cfg, "pad_token_id", None))
    B, S = inputs_embeds.shape[:2]
    if position_ids is None:
        position_ids = torch.arange(S)[None, :].expand(B, S)
    if token_type_ids is None:
        token_type_ids = torch.zeros(B, S, dtype=torch.long)
    e = (inputs_embeds
         + sd[prefix + "token_type_embeddings.weight"][token_type_ids]
         + sd[prefix + "position_embeddings.weight"][position_ids])
    y = layer_norm(e, sd[prefix + "LayerNorm.weight"], sd[prefix + "LayerNorm.bias"],
                   cfg.layer_norm_eps)
    return _drop(y, masks, "emb")


def _drop(x: Tensor, masks: Optional[Dict[str, Tensor]], key: str) -> Tensor:
    """nn.Dropout in training mode with the mask made explicit: `masks[key]` holds 0 or 1/(1-p) per element (absent key
    = identity).  The reference draws its masks from torch's generator; which elements fall is not part of the
    algorithm, so parity under dropout is "same function of (input, mask)"."""
    if masks is None or key not in masks:
        return x
    return x * masks[key].reshape(x.shape).to(x.dtype)


def attention_core(q: Tensor, k: Tensor, v: Tensor, n_heads: int,
                   add_mask: Optional[Tensor], prob_mask: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """softmax(Q K^T / sqrt(d) + mask) V with head split/merge.

    q: [B,Sq,H]; k,v: [B,Sk,H].  Returns (context [B,Sq,H], probs [B,h,Sq,Sk]).
    bert_model.py:252-257 (transpose_for_scores), :309 (QK^T), :328 (/sqrt(d)),
    :329-331 (+mask), :334 (softmax), :343 (P V), :345-347 (merge heads).
    """
    B, Sq, H = q.shape
    Sk = k.shape[1]
    d = H // n_heads
    qh = q.view(B, Sq, n_heads, d).permute(0, 2, 1, 3)
    kh = k.view(B, Sk, n_heads, d).permute(0, 2, 1, 3)
    vh = v.view(B, Sk, n_heads, d).permute(0, 2, 1, 3)
    scores = (qh @ kh.transpose(-1, -2)) / math.sqrt(d)
    if add_mask is not None:
        scores = scores + add_mask
    probs = torch.softmax(scores, dim=-1)
    dropped = probs if prob_mask is None else probs * prob_mask.to(probs.dtype)       # bert_model.py:338
    ctx = (dropped @ vh).permute(0, 2, 1, 3).reshape(B, Sq, H)
    return ctx, probs


def self_attention_block(sd, p: str, cfg: OracleConfig, x: Tensor, add_mask,
                         kv_states: Optional[Tensor] = None, masks: Optional[Dict[str, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """BertAttention = BertSelfAttention + BertSelfOutput.

    `kv_states` (cross-attention, bert_model.py:283-286) makes K/V come from
    another tensor (whose last dim may differ: `ce_kv_hidden_size`).
    """
    src = x if kv_states is None else kv_states
    q = linear(x, sd[p + "self.query.weight"], sd[p + "self.query.bias"])
    k = linear(src, sd[p + "self.key.weight"], sd[p + "self.key.bias"])
    v = linear(src, sd[p + "self.value.weight"], sd[p + "self.value.bias"])
    ctx, probs = attention_core(q, k, v, cfg.num_attention_heads, add_mask, None if masks is None else masks.get(p + "probs"))
    y = _drop(linear(ctx, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"]), masks, p + "out")       # :373
    y = layer_norm(y + x, sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"],
                   cfg.layer_norm_eps)
    return y, probs


def ffn_block(sd, p: str, cfg: OracleConfig, x: Tensor, masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """BertIntermediate + BertOutput — bert_model.py:436-439, 449-453."""
    h = gelu_erf(linear(x, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
    y = _drop(linear(h, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"]), masks, p + "ffn_out")     # :451
    return layer_norm(y + x, sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"],
                      cfg.layer_norm_eps)


def bert_layer(sd, p: str, cfg: OracleConfig, x: Tensor, add_mask, masks: Optional[Dict[str, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """One BertLayer / BertSelfAttnLayer — bert_model.py:518-553.  Dropout multipliers (training mode) are looked up in
    `masks` under p+"attention.probs", p+"attention.out", p+"ffn_out"."""
    a, probs = self_attention_block(sd, p + "attention.", cfg, x, add_mask, masks=masks)
    return ffn_block(sd, p, cfg, a, masks), probs


def bert_cross_layer(sd, p: str, cfg: OracleConfig, x: Tensor, kv: Tensor,
                     self_mask, cross_mask, do_ffn: bool = True) -> Tensor:
    """BertCrossLayer — bert_model.py:456-515: self-attn block, cross-attn block
    (K/V projected from `kv`, width ce_kv_hidden_size), then FFN."""
    a, _ = self_attention_block(sd, p + "attention.", cfg, x, self_mask)
    c, _ = self_attention_block(sd, p + "crossattention.", cfg, a, cross_mask, kv_states=kv)
    return ffn_block(sd, p, cfg, c) if do_ffn else c


def pooler(sd, h: Tensor, prefix: str = "pooler.") -> Tensor:
    """tanh(Linear(h[:,0])) — bert_model.py:689-701."""
    return torch.tanh(linear(h[:, 0], sd[prefix + "dense.weight"], sd[prefix + "dense.bias"]))


@dataclass
class BertOracleOutput:
    last_hidden_state: Tensor
    pooler_output: Optional[Tensor]
    hidden_states: List[Tensor]      # L+1 entries: embeddings output, then each layer
    attentions: List[Tensor]         # L entries [B,h,S,S]


def bert_model(sd: Dict[str, Tensor], cfg: OracleConfig, input_ids: Optional[Tensor] = None,
               attention_mask: Optional[Tensor] = None, token_type_ids: Optional[Tensor] = None,
               position_ids: Optional[Tensor] = None, inputs_embeds: Optional[Tensor] = None,
               mask_fill: Optional[float] = None, masks: Optional[Dict[str, Tensor]] = None) -> BertOracleOutput:
    """BertModel.forward; eval mode (dropouts are identity) unless explicit dropout multipliers are given in `masks`."""
    x = embeddings(sd, cfg, input_ids, token_type_ids, position_ids, inputs_embeds, masks=masks)
    add_mask = additive_key_mask(attention_mask, x.dtype, mask_fill)
    hs, atts = [x], []
    for i in range(cfg.num_hidden_layers):
        x, probs = bert_layer(sd, f"encoder.layer.{i}.", cfg, x, add_mask, masks)
        hs.append(x)
        atts.append(probs)
    pooled = pooler(sd, x) if "pooler.dense.weight" in sd else None
    return BertOracleOutput(x, pooled, hs, atts)


# ----------------------------------------------------------------------------
# heads / losses / poolers that sit directly on the encoder output
# ----------------------------------------------------------------------------

def token_cls_logits(h: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """LossCalculator.classifier (loss_calculator.py:17,42): Linear(H -> num_labels)."""
    return linear(h, w, b)


def cross_entropy(logits: Tensor, labels: Tensor, ignore_index: int = -100,
                  class_weight: Optional[Tensor] = None) -> Tensor:
    """torch CrossEntropyLoss(weight, ignore_index=-100, reduction='mean')
    (utils.py:173-182 with gamma == 0), written out."""
    C = logits.shape[-1]
    lg = logits.reshape(-1, C)
    lb = labels.reshape(-1)
    keep = lb != ignore_index
    lsm = lg - torch.logsumexp(lg, dim=-1, keepdim=True)
    safe = torch.where(keep, lb, torch.zeros_like(lb))
    nll = -lsm.gather(1, safe[:, None])[:, 0]
    w = torch.ones_like(nll) if class_weight is None else class_weight[safe]
    w = torch.where(keep, w, torch.zeros_like(w))
    return (nll * w).sum() / w.sum()


def boundary_argmax(logits: Tensor) -> Tensor:
    """np.argmax(logits, axis=-1) — ts_sentence_seq_labeling.py:1032,1143."""
    return logits.argmax(dim=-1)


def ditto_pool(out: BertOracleOutput, attention_mask: Tensor, layer: int, head: int,
               pooler_type: str = "att_first_last") -> Tensor:
    """ditto/evaluation_ditto.py:125-155: importance = diag of one attention head;
    weighted sum of (first+last)/2 hidden states (or last only)."""
    diag = torch.diagonal(out.attentions[layer][:, head], dim1=1, dim2=2)      # [B,S]
    m = attention_mask.to(diag.dtype)
    if pooler_type == "att_first_last":
        h = (out.hidden_states[0] + out.hidden_states[-1]) / 2.0
    elif pooler_type == "att_last":
        h = out.hidden_states[-1]
    else:
        raise ValueError(pooler_type)
    return (h * m[:, :, None] * diag[:, :, None]).sum(1)


# ----------------------------------------------------------------------------
# training-step oracle: scalar loss + autograd gradients (fp32, CPU)
# ----------------------------------------------------------------------------

def topicseg_loss(sd: Dict[str, Tensor], cfg: OracleConfig, cls_w: Tensor, cls_b: Tensor,
                  input_ids: Tensor, attention_mask: Tensor, token_type_ids: Tensor,
                  labels: Tensor, masks: Optional[Dict[str, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """Encoder -> Linear(H->2) -> CE(ignore -100): the `ts_score_predictor == "lt"`
    path of LossCalculator with CSSL/TSSP weights 0 and dropout off.  Used for
    gradient parity (the wrapper itself cannot run backward on CPU: SURVEY §8c trap 1)."""
    out = bert_model(sd, cfg, input_ids, attention_mask, token_type_ids, masks=masks)
    logits = token_cls_logits(_drop(out.last_hidden_state, masks, "head"), cls_w, cls_b)       # bert_for_ts.py:66-67
    return cross_entropy(logits, labels), logits


def random_state_dict(cfg: OracleConfig, seed: int = 0, std: float = 0.02,
                      with_pooler: bool = True) -> Dict[str, Tensor]:
    """HF `_init_weights`-style init (N(0, 0.02); LN gamma 1 beta 0; biases 0) with
    small random biases / LN params added so that bias- and affine-handling bugs
    cannot hide behind zeros."""
    g = torch.Generator().manual_seed(seed)
    H, I = cfg.hidden_size, cfg.intermediate_size

    def n(*shape, s=std):
        return torch.randn(*shape, generator=g) * s

    sd = {
        "embeddings.word_embeddings.weight": n(cfg.vocab_size, H),
        "embeddings.position_embeddings.weight": n(cfg.max_position_embeddings, H),
        "embeddings.token_type_embeddings.weight": n(cfg.type_vocab_size, H),
        "embeddings.LayerNorm.weight": 1.0 + n(H, s=0.05),
        "embeddings.LayerNorm.bias": n(H, s=0.05),
    }
    for i in range(cfg.num_hidden_layers):
        p = f"encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            sd[p + f"attention.self.{nm}.weight"] = n(H, H)
            sd[p + f"attention.self.{nm}.bias"] = n(H)
        sd[p + "attention.output.dense.weight"] = n(H, H)
        sd[p + "attention.output.dense.bias"] = n(H)
        sd[p + "attention.output.LayerNorm.weight"] = 1.0 + n(H, s=0.05)
        sd[p + "attention.output.LayerNorm.bias"] = n(H, s=0.05)
        sd[p + "intermediate.dense.weight"] = n(I, H)
        sd[p + "intermediate.dense.bias"] = n(I)
        sd[p + "output.dense.weight"] = n(H, I)
        sd[p + "output.dense.bias"] = n(H)
        sd[p + "output.LayerNorm.weight"] = 1.0 + n(H, s=0.05)
        sd[p + "output.LayerNorm.bias"] = n(H, s=0.05)
    if with_pooler:
        sd["pooler.dense.weight"] = n(H, H)
        sd["pooler.dense.bias"] = n(H)
    return sd
