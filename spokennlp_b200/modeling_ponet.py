"""Drop-in `PoNetModel` for alimeeting4mug's topic-segmentation script, backed by libb200enc.so.

Boundary: `modelscope.models.nlp.ponet.PoNetModel` as constructed and called at
alimeeting4mug/src/models/modeling_ponet.py:41 (`PoNetModel(config, add_pooling_layer=False)`) and :68-79
(`self.ponet(input_ids, attention_mask=..., token_type_ids=..., segment_ids=..., position_ids=..., head_mask=...,
inputs_embeds=..., output_attentions=..., output_hidden_states=..., return_dict=...)`).  modelscope is not installable
here, so the layer arithmetic follows the restatement in oracle/ponet_oracle.py (**parity unpinned**) and
the parameter names follow the public PoNet implementation (`attention.self.dense_{q,k,o,segment,local}`).

Per layer: ONE packed [5H,H] tcgen05 GEMM produces Q|K|O|Sg|Lc; the pooling mixer (global softmax-pooled vector per
head, segment max, local max-3, fuse) runs as coalesced HBM-bound kernels (`b200_ponet_mix_fwd`); the output dense +
residual + LayerNorm and the feed-forward block are the same kernels as BERT.  Forward and backward (fine-tuning) both run
in the library; gradients reach every parameter through one autograd node.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn
from transformers import PretrainedConfig
from transformers.modeling_outputs import BaseModelOutputWithPoolingAndCrossAttentions

from . import ops
from .blocks import FfnWeights, ffn_block_bwd, ffn_block_fwd
from .engine import EMB_NAMES, DropPlan, FlatParams, next_drop_seed
from .lib import B200Error
from .modeling_bert import BertEmbeddings, BertIntermediate, BertOutput, BertPooler, BertSelfOutput

F16, F32 = torch.float16, torch.float32
_PROJ = ("q", "k", "o", "segment", "local")


class PoNetConfig(PretrainedConfig):
    model_type = "ponet"

    def __init__(self, vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=512,
                 type_vocab_size=2, initializer_range=0.02, layer_norm_eps=1e-12, pad_token_id=0, **kwargs):
        super().__init__(pad_token_id=pad_token_id, **kwargs)
        self.vocab_size, self.hidden_size, self.num_hidden_layers = vocab_size, hidden_size, num_hidden_layers
        self.num_attention_heads, self.intermediate_size, self.hidden_act = num_attention_heads, intermediate_size, hidden_act
        self.hidden_dropout_prob, self.attention_probs_dropout_prob = hidden_dropout_prob, attention_probs_dropout_prob
        self.max_position_embeddings, self.type_vocab_size = max_position_embeddings, type_vocab_size
        self.initializer_range, self.layer_norm_eps = initializer_range, layer_norm_eps


class PoNetSelfAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        for n in _PROJ:
            setattr(self, f"dense_{n}", nn.Linear(config.hidden_size, config.hidden_size))


class PoNetAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = PoNetSelfAttention(config)
        self.output = BertSelfOutput(config)


class PoNetLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = PoNetAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)


class PoNetEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([PoNetLayer(config) for _ in range(config.num_hidden_layers)])


def ponet_layer_names(i: int) -> List[str]:
    p = f"encoder.layer.{i}."
    return ([p + f"attention.self.dense_{n}.weight" for n in _PROJ] + [p + f"attention.self.dense_{n}.bias" for n in _PROJ] +
            [p + "attention.output.dense.weight", p + "attention.output.dense.bias", p + "attention.output.LayerNorm.weight",
             p + "attention.output.LayerNorm.bias", p + "intermediate.dense.weight", p + "intermediate.dense.bias",
             p + "output.dense.weight", p + "output.dense.bias", p + "output.LayerNorm.weight", p + "output.LayerNorm.bias"])


class _PoNetFn(torch.autograd.Function):
    """Embeddings + L PoNet layers as one autograd node (same bridge as modeling_bert._EncoderFn)."""

    @staticmethod
    def forward(ctx, model, ids, tt, pos, key_bias, seg, B, S, want_hidden, drop, save, *params):
        # `save` (= gradients will be asked for) is decided by the module, where the grad mode is visible; `drop`: DropPlan or
        # None — hidden dropout after the embeddings, the attention output dense and the FFN output dense, as in BERT (PoNet's
        # pooling mixer has no probabilities to drop).
        f: FlatParams = model._flat
        cfg = model.config
        H, heads, eps, dev = cfg.hidden_size, cfg.num_attention_heads, float(cfg.layer_norm_eps), ids.device
        M, nseg = B * S, S + 2
        drop_emb = drop.at(DropPlan.EMB, drop.p_hidden) if drop is not None else None
        x32 = torch.empty(M, H, dtype=F32, device=dev)
        x16 = ops.embed_ln_fwd(ids, tt, pos, None, f.view32(EMB_NAMES[0]), f.view32(EMB_NAMES[1]), f.view32(EMB_NAMES[2]),
                               f.view32(EMB_NAMES[3]), f.view32(EMB_NAMES[4]), eps, M, S, H, y32=x32, drop=drop_emb)
        hiddens = [x32.view(B, S, H)] if want_hidden else []
        saved = []
        for i in range(cfg.num_hidden_layers):
            n = ponet_layer_names(i)
            proj = torch.empty(M, 5 * H, dtype=F16, device=dev)
            ops.gemm(x16, f.view16(n[0], tuple(n[1:5])), proj, epilogue=ops.EPI_BIAS, bias=f.view32(n[5], tuple(n[6:10])))
            mix = torch.empty(M, H, dtype=F16, device=dev)
            ws = ops.ponet_mix_fwd(proj, seg, mix, B, S, heads, nseg, key_bias=key_bias)
            pre = torch.empty(M, H, dtype=F32, device=dev)
            d_ao = drop.layer(i, DropPlan.ATTN_OUT) if drop is not None else None
            ops.gemm(mix, f.view16(n[10]), pre, epilogue=ops.EPI_BIAS_RES32, bias=f.view32(n[11]), aux=x32, drop=d_ao)
            mean = torch.empty(M, dtype=F32, device=dev) if save else None
            rstd = torch.empty(M, dtype=F32, device=dev) if save else None
            a32 = torch.empty(M, H, dtype=F32, device=dev)
            a16 = ops.layernorm_fwd(pre, f.view32(n[12]), f.view32(n[13]), eps, y32=a32, mean=mean, rstd=rstd)
            y16, y32, svf = ffn_block_fwd(_ffn_views(f, n, "p"), a16, a32, eps, save=save,
                                          drop_hidden=drop.layer(i, DropPlan.FFN_OUT) if drop is not None else None)
            if save:
                saved.append((x16, proj, ws, mix, pre, mean, rstd, svf, d_ao))
            x16, x32 = y16, y32
            if want_hidden:
                hiddens.append(x32.view(B, S, H))
        ctx.model, ctx.saved, ctx.meta = model, (saved if save else None), (ids, tt, pos, key_bias, seg, B, S, drop_emb)
        outs = [x32.view(B, S, H)] + (hiddens if want_hidden else [])
        ctx.mark_non_differentiable(*outs[1:])
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_last, *unused):
        model, saved = ctx.model, ctx.saved
        if saved is None:
            raise B200Error("backward through a forward that ran without grad")
        ids, tt, pos, key_bias, seg, B, S, drop_emb = ctx.meta
        f: FlatParams = model._flat
        cfg = model.config
        H, heads, eps, dev = cfg.hidden_size, cfg.num_attention_heads, float(cfg.layer_norm_eps), g_last.device
        M, nseg = B * S, S + 2
        dy = torch.empty(M, H, dtype=F16, device=dev)
        scale = torch.empty(2, dtype=F32, device=dev)
        slot = torch.empty(1, dtype=torch.int32, device=dev)
        ops.scale_cast_grad(g_last.contiguous().to(F32).view(-1), dy.view(-1), scale, slot, target=1024.0)
        inv = scale[1:2]
        keep, f.grad32 = f.grad32, torch.zeros_like(f.flat32)
        try:
            for i in reversed(range(cfg.num_hidden_layers)):
                n = ponet_layer_names(i)
                x16, proj, ws, mix, pre, mean, rstd, svf, d_ao = saved[i]
                saved[i] = None
                d_a = ffn_block_bwd(_ffn_views(f, n, "p"), _ffn_views(f, n, "g"), svf, dy, inv)
                d_pre = torch.empty(M, H, dtype=F16, device=dev)
                d_den = d_pre                        # gradient wrt the dense output = d_pre x the regenerated dropout mask
                if d_ao is not None and d_ao.p > 0.0:
                    d_den = torch.empty_like(d_pre)
                    ops.layernorm_bwd(d_a, pre, mean, rstd, f.view32(n[12]), d_pre, f.viewg(n[12]), f.viewg(n[13]), dbias=f.viewg(n[11]),
                                      alpha=inv, dx_drop=d_den, drop=d_ao)
                else:
                    ops.layernorm_bwd(d_a, pre, mean, rstd, f.view32(n[12]), d_pre, f.viewg(n[12]), f.viewg(n[13]), dbias=f.viewg(n[11]), alpha=inv)
                ops.gemm(d_den, mix, f.viewg(n[10]), a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv, k_splits=ops.wgrad_splits(H, H, M))
                dmix = torch.empty(M, H, dtype=F16, device=dev)
                ops.gemm(d_den, f.view16(n[10]), dmix, b_layout=1)
                dproj = torch.empty(M, 5 * H, dtype=F16, device=dev)
                ops.ponet_mix_bwd(proj, dmix, seg, ws, dproj, B, S, heads, nseg, key_bias=key_bias)
                ops.colsum(dproj, f.viewg(n[5], tuple(n[6:10])), inv)
                ops.gemm(dproj, x16, f.viewg(n[0], tuple(n[1:5])), a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv,
                         k_splits=ops.wgrad_splits(5 * H, H, M))
                dy = torch.empty(M, H, dtype=F16, device=dev)
                ops.gemm(dproj, f.view16(n[0], tuple(n[1:5])), dy, b_layout=1, epilogue=ops.EPI_ADD, aux=d_pre)
            ops.embed_ln_bwd(dy, None, ids, tt, pos, f.view32(EMB_NAMES[0]), f.view32(EMB_NAMES[1]), f.view32(EMB_NAMES[2]),
                             f.view32(EMB_NAMES[3]), f.viewg(EMB_NAMES[0]), f.viewg(EMB_NAMES[1]), f.viewg(EMB_NAMES[2]),
                             f.viewg(EMB_NAMES[3]), f.viewg(EMB_NAMES[4]), inv, eps, M, S, H, drop=drop_emb,
                             pad_id=model.embeddings.word_embeddings.padding_idx)
            grads = tuple(f.viewg(nm) if f.params[nm].requires_grad else None for nm in f.names)
        finally:
            f.grad32 = keep
        ctx.saved = None
        return (None,) * 11 + grads


def _ffn_views(f: FlatParams, n, kind: str) -> FfnWeights:
    w = {"p": f.view16, "g": f.viewg}[kind]
    s = {"p": f.view32, "g": f.viewg}[kind]
    return FfnWeights(w1=w(n[14]), bf1=s(n[15]), w2=w(n[16]), bf2=s(n[17]), g=s(n[18]), b=s(n[19]))


class PoNetModel(nn.Module):
    def __init__(self, config, add_pooling_layer: bool = True):
        super().__init__()
        self.config = config
        if config.hidden_size != 64 * config.num_attention_heads:
            raise B200Error("B200 PoNetModel requires head_dim == 64")
        self.embeddings = BertEmbeddings(config)
        self.encoder = PoNetEncoder(config)
        self.pooler = BertPooler(config) if add_pooling_layer else None
        std = getattr(config, "initializer_range", 0.02)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                m.weight.data.normal_(mean=0.0, std=std)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()
        self._flat: Optional[FlatParams] = None

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def set_input_embeddings(self, value):
        self.embeddings.word_embeddings = value

    def _packed(self, device) -> FlatParams:
        f = self._flat
        own = dict(self.named_parameters())
        names = list(EMB_NAMES)
        for i in range(self.config.num_hidden_layers):
            names += ponet_layer_names(i)
        named = [(n, own[n]) for n in names]
        if f is not None and f.intact(named):        # also notices a replaced Parameter object (resize_token_embeddings)
            f.sync_half()
            return f
        if torch.device(device).type != "cuda":
            raise B200Error("B200 PoNetModel runs on CUDA devices only (no CPU fallback)")
        self._flat = FlatParams(named, device)
        return self._flat

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, segment_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, output_attentions=None, output_hidden_states=None, return_dict=None, **kwargs):
        cfg = self.config
        if head_mask is not None or output_attentions:
            raise B200Error("head_mask / output_attentions are not supported by the B200 PoNetModel")
        if input_ids is None:
            raise B200Error("B200 PoNetModel needs input_ids")
        if not input_ids.is_cuda:
            raise B200Error("B200 PoNetModel got CPU tensors: there is no CPU fallback")
        if segment_ids is None:
            raise B200Error("PoNetModel.forward needs segment_ids (modeling_ponet.py:72)")
        return_dict = True if return_dict is None else return_dict
        B, S = input_ids.shape
        f = self._packed(input_ids.device)
        if position_ids is None:
            position_ids = self.embeddings.position_ids[:, :S].expand(B, S)     # honours the driver's in-place 4096 tiling
        pos = position_ids.expand(B, S).contiguous().view(-1)
        tt = token_type_ids.contiguous().view(-1) if token_type_ids is not None else None
        key_bias = None
        if attention_mask is not None:
            key_bias, _ = ops.mask_to_bias(attention_mask)
        seg = segment_ids.contiguous().to(torch.int64)
        params = [f.params[n] for n in f.names]
        drop = None
        p_hidden = float(getattr(cfg, "hidden_dropout_prob", 0.0) or 0.0)
        if self.training and p_hidden > 0.0:
            drop = DropPlan(next_drop_seed(self, input_ids.device), p_hidden, 0.0)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        outs = _PoNetFn.apply(self, input_ids.contiguous().view(-1), tt, pos, key_bias, seg, B, S, bool(output_hidden_states), drop,
                              need_grad, *params)
        seq = outs[0]
        hs = None
        if output_hidden_states:
            hs = tuple(outs[1:-1]) + (seq,)
        pooled = self.pooler(seq) if self.pooler is not None else None
        if not return_dict:
            return (seq, pooled) + ((hs,) if hs is not None else ())
        return BaseModelOutputWithPoolingAndCrossAttentions(last_hidden_state=seq, pooler_output=pooled, hidden_states=hs, attentions=None)
