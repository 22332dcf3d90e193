"""Drop-in `BertModel` for SpokenNLP's scripts, backed by libb200enc.so (sm_100a).

Boundary being matched: `transformers.models.bert.modeling_bert.BertModel` as called at
  emnlp2023-topic_segmentation/src/models/bert_for_ts.py:55-65,70-80   (positional input_ids, return_dict=False)
  mmvts/src/models/text_encoder/text_encoder.py:61-71
  ditto/evaluation_ditto.py:121  (output_hidden_states=True, output_attentions=True, return_dict=True)
Same constructor, same forward signature, same `state_dict` keys, same tuple / ModelOutput returns.  The module tree
below only HOLDS parameters (so HF checkpoints, `resize_token_embeddings`, optimizers and DDP see the usual names);
all arithmetic runs in the CUDA library through `EncoderEngine`.  There is no CPU / PyTorch fallback: calling the
model on CPU tensors raises.
"""
from __future__ import annotations

import warnings
from typing import Optional

import torch
from torch import nn
from transformers.modeling_outputs import BaseModelOutputWithPoolingAndCrossAttentions
from transformers.models.bert.modeling_bert import BertPreTrainedModel

from . import ops
from .engine import EMB_NAMES, DropPlan, EncoderEngine, FlatParams, layer_param_names, next_drop_seed
from .lib import B200Error


# ---------------------------------------------------------------------------- parameter holders (HF names)
class BertEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)), persistent=False)
        self.register_buffer("token_type_ids", torch.zeros((1, config.max_position_embeddings), dtype=torch.long),
                             persistent=False)


class BertSelfAttention(nn.Module):
    def __init__(self, config, kv_hidden_size: Optional[int] = None):
        super().__init__()
        kv = kv_hidden_size or config.hidden_size
        self.query = nn.Linear(config.hidden_size, config.hidden_size)
        self.key = nn.Linear(kv, config.hidden_size)
        self.value = nn.Linear(kv, config.hidden_size)


class BertSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertAttention(nn.Module):
    def __init__(self, config, kv_hidden_size: Optional[int] = None):
        super().__init__()
        self.self = BertSelfAttention(config, kv_hidden_size)
        self.output = BertSelfOutput(config)


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class BertOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = BertAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)


class BertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(config) for _ in range(config.num_hidden_layers)])


class BertPooler(nn.Module):
    """tanh(Linear(h[:,0])) — bert_model.py:689-701.  A [B,H]x[H,H] product: host-side glue in fp32, not a hot op
    (the topic-segmentation wrappers drop the pooler: bert_for_ts.py:20)."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)

    def forward(self, hidden_states):
        return torch.tanh(self.dense(hidden_states[:, 0]))


# ---------------------------------------------------------------------------- autograd bridge
class _EncoderFn(torch.autograd.Function):
    """Whole embeddings+encoder stack as one autograd node.  Outputs: last hidden state (fp32, differentiable) followed
    by optional per-layer hidden states and attention probabilities (returned detached)."""

    @staticmethod
    def forward(ctx, model, ids, tt, pos, inputs_embeds, key_bias, kv_len, B, S, want_hidden, want_probs, drop, need_grad, *params):
        # `need_grad` is decided by the module's forward, where the grad MODE is visible: ctx.needs_input_grad only mirrors
        # requires_grad of the inputs and stays True under torch.no_grad() / in eval, which used to make every inference
        # forward save all L layers' activations.
        eng: EncoderEngine = model._engine
        x16, x32, saved, hiddens, probs = eng.forward(ids, tt, pos, inputs_embeds, key_bias, kv_len, B, S, save=need_grad,
                                                 want_hidden=want_hidden, want_probs=want_probs, drop=drop)
        ctx.model, ctx.saved, ctx.n_params = model, saved, len(params)
        H = eng.H

        outs = [x32.view(B, S, H)]             # the fp32 copy the last LayerNorm wrote (no extra cast pass)
        if want_hidden:
            outs += [h.view(B, S, H) for h in hiddens]
        if want_probs:
            outs += probs
        ctx.mark_non_differentiable(*outs[1:])
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_last, *unused):
        model, saved = ctx.model, ctx.saved
        if saved is None:
            raise B200Error("backward through a forward that ran without grad")
        eng: EncoderEngine = model._engine
        flat = eng.flat
        dev = g_last.device
        g_last = g_last.contiguous().float()
        dy = torch.empty(saved.B * saved.S, eng.H, dtype=torch.float16, device=dev)
        scale = torch.empty(2, dtype=torch.float32, device=dev)
        slot = torch.empty(1, dtype=torch.int32, device=dev)
        ops.scale_cast_grad(g_last.view(-1), dy.view(-1), scale, slot, target=1024.0)
        keep, flat.grad32 = flat.grad32, torch.zeros_like(flat.flat32)       # fresh buffer: autograd owns the result
        try:
            d_emb = eng.backward(saved, dy, scale[1:2], want_d_inputs_embeds=ctx.needs_input_grad[4])
            grads = tuple(flat.viewg(n) if flat.params[n].requires_grad else None for n in flat.names)
        finally:
            flat.grad32 = keep
        ctx.saved = None
        if d_emb is not None:
            d_emb = d_emb.view(saved.B, saved.S, eng.H)
        return (None,) * 4 + (d_emb,) + (None,) * 8 + grads


# ---------------------------------------------------------------------------- the model
class BertModel(BertPreTrainedModel):
    """B200-native `BertModel`.  HF-identical constructor / forward signature / state_dict."""

    def __init__(self, config, add_pooling_layer: bool = True):
        super().__init__(config)
        self.config = config
        if config.hidden_size != 64 * config.num_attention_heads:
            raise B200Error("B200 BertModel requires head_dim == 64 (hidden_size == 64 * num_attention_heads)")
        if getattr(config, "hidden_act", "gelu") != "gelu":
            raise B200Error("B200 BertModel implements hidden_act='gelu' (erf) only")
        if getattr(config, "position_embedding_type", "absolute") not in (None, "absolute"):
            raise B200Error("B200 BertModel implements absolute position embeddings only")
        self.embeddings = BertEmbeddings(config)
        self.encoder = BertEncoder(config)
        self.pooler = BertPooler(config) if add_pooling_layer else None
        self._engine: Optional[EncoderEngine] = None
        self.post_init()

    # HF plumbing used by the reference drivers (ts_sentence_seq_labeling.py:284, main_multimodal.py:291)
    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def set_input_embeddings(self, value):
        self.embeddings.word_embeddings = value

    # ---- packing ---------------------------------------------------------------------------------------------------
    def _hot_named_params(self):
        own = dict(self.named_parameters())
        names = list(EMB_NAMES)
        for i in range(self.config.num_hidden_layers):
            names += layer_param_names(i)
        return [(n, own[n]) for n in names]

    def b200_engine(self, device=None, named=None) -> EncoderEngine:
        """(Re)build the packed parameter buffers if the module's parameters moved or were replaced (``.to()``, resize, load)."""
        eng = self._engine
        named = named if named is not None else self._hot_named_params()
        if eng is not None and eng.flat.intact(named):
            return eng
        device = device or named[0][1].device
        if torch.device(device).type != "cuda":
            raise B200Error("B200 BertModel runs on CUDA devices only (no CPU fallback): move the model with .cuda()")
        flat = FlatParams(named, device)
        c = self.config
        self._engine = EncoderEngine(flat, c.hidden_size, c.num_attention_heads, c.intermediate_size, c.num_hidden_layers,
                                     float(c.layer_norm_eps), pad_id=self.embeddings.word_embeddings.padding_idx)
        return self._engine

    # ---- forward ---------------------------------------------------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, output_attentions=None, output_hidden_states=None, return_dict=None, **kwargs):
        cfg = self.config
        output_attentions = cfg.output_attentions if output_attentions is None else output_attentions
        output_hidden_states = cfg.output_hidden_states if output_hidden_states is None else output_hidden_states
        return_dict = getattr(cfg, "use_return_dict", True) if return_dict is None else return_dict
        if head_mask is not None:
            raise B200Error("head_mask is not supported by the B200 encoder (every reference call site passes None)")
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
        src = input_ids if input_ids is not None else inputs_embeds
        if not src.is_cuda:
            raise B200Error("B200 BertModel got CPU tensors: there is no CPU fallback")
        B, S = src.shape[0], src.shape[1]
        if S > cfg.max_position_embeddings and position_ids is None:
            raise ValueError(f"sequence length {S} > max_position_embeddings {cfg.max_position_embeddings}")
        named = self._hot_named_params()
        eng = self.b200_engine(src.device, named)
        drop = None
        if self.training and (cfg.hidden_dropout_prob > 0 or cfg.attention_probs_dropout_prob > 0):
            # one base seed per forward (engine.next_drop_seed: first value from torch's CPU generator, so torch.manual_seed governs it as
            # it does nn.Dropout in the reference; advanced on the device); the tensor lives in the saved state until this forward's
            # backward has run
            drop = DropPlan(next_drop_seed(self, src.device), cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob)

        ids = input_ids.contiguous().view(-1) if input_ids is not None else None
        emb = inputs_embeds.contiguous().float() if inputs_embeds is not None else None       # [B,S,H]; the kernels index it as [B*S,H]
        tt = token_type_ids.contiguous().view(-1) if token_type_ids is not None else None
        if position_ids is None:
            # honour an in-place edited buffer (ponet_topic_segmentation.py:471-482 rewrites embeddings.position_ids)
            position_ids = self.embeddings.position_ids[:, :S].expand(B, S)
        pos = position_ids.expand(B, S).contiguous().view(-1)
        key_bias = kv_len = None
        if attention_mask is not None:
            if attention_mask.dim() != 2:
                raise B200Error("attention_mask must be [batch, seq] (key padding mask)")
            key_bias, kv_len = ops.mask_to_bias(attention_mask)

        params = [p for _, p in named]
        need_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or (emb is not None and emb.requires_grad))
        outs = _EncoderFn.apply(self, ids, tt, pos, emb, key_bias, kv_len, B, S, bool(output_hidden_states),
                                bool(output_attentions), drop, need_grad, *params)
        seq = outs[0]
        k = 1
        hidden_states = attentions = None
        if output_hidden_states:
            hidden_states = tuple(outs[k:k + cfg.num_hidden_layers + 1])
            k += cfg.num_hidden_layers + 1
            hidden_states = hidden_states[:-1] + (seq,)
        if output_attentions:
            attentions = tuple(outs[k:k + cfg.num_hidden_layers])
        pooled = self.pooler(seq) if self.pooler is not None else None
        if not return_dict:
            out = (seq, pooled)
            if output_hidden_states:
                out = out + (hidden_states,)
            if output_attentions:
                out = out + (attentions,)
            return out
        return BaseModelOutputWithPoolingAndCrossAttentions(last_hidden_state=seq, pooler_output=pooled,
                                                            hidden_states=hidden_states, attentions=attentions)


def patch_transformers() -> None:
    """Make `from transformers.models.bert.modeling_bert import BertModel` (what the reference wrappers do:
    bert_for_ts.py:7, text_encoder.py:19) resolve to the B200 implementation.  Call before importing them."""
    import transformers
    import transformers.models.bert.modeling_bert as mb
    mb.BertModel = BertModel
    transformers.BertModel = BertModel
