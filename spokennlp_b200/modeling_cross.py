"""Drop-in cross-modal encoder layers of mmvts, backed by libb200enc.so.

Boundary — same class names, constructor arguments, `state_dict` keys and forward
signatures as the reference:
  * `BertSelfAttnLayer(bert_config, config)`             mmvts/src/models/cross_encoder/bert_model.py:518-553
  * `BertCrossLayer(config, ce_kv_hidden_size)`          mmvts/src/models/cross_encoder/bert_model.py:456-515
  * `MergeAttentionEncoder(config)`                      mmvts/src/models/cross_encoder/ma_encoder.py:10-71
  * `CoAttentionEncoder(config)`                         mmvts/src/models/cross_encoder/ca_encoder.py:13-77
  * `LinearProjector(config)`                            mmvts/src/models/projector/linear_projector.py:5-30
Inputs / outputs are the reference's fp32 `[B, N, H]` features and `[B,1,1,N]` additive masks (0 / -1e6); gradients flow
to the inputs (the text encoder upstream) and to every parameter.  All arithmetic runs in the CUDA library; `torch.cat`
/ `torch.chunk` of the encoders are the reference's own data movement and stay in torch.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import ops
from .blocks import (AttnWeights, FfnWeights, attn_block_bwd, attn_block_fwd, ffn_block_bwd, ffn_block_fwd)
from .engine import DropPlan, FlatParams, next_drop_seed
from .lib import B200Error
from .modeling_bert import BertAttention, BertIntermediate, BertOutput

Tensor = torch.Tensor
F16, F32 = torch.float16, torch.float32

_ATTN = ["self.query.weight", "self.key.weight", "self.value.weight", "self.query.bias", "self.key.bias", "self.value.bias",
         "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias"]
_FFN = ["intermediate.dense.weight", "intermediate.dense.bias", "output.dense.weight", "output.dense.bias",
        "output.LayerNorm.weight", "output.LayerNorm.bias"]


def _mask_to_key_bias(mask: Optional[Tensor], B: int, N: int) -> Optional[Tensor]:
    """[B,1,1,N] additive mask (ma_encoder.py:55-56, ca_encoder.py:48-49) or [B,N] -> fp32 [B,N] additive key bias."""
    if mask is None:
        return None
    if mask.numel() != B * N:
        raise B200Error(f"attention mask with {mask.numel()} elements does not match [B={B}, N={N}] keys")
    return mask.reshape(B, N).to(F32).contiguous()


def _to16(x32: Tensor) -> Tensor:
    out = torch.empty(x32.shape, dtype=F16, device=x32.device)
    return ops.cast_f32_to_f16(x32, out)


class _Packed(nn.Module):
    """Owns a flat fp32/fp16 copy of its parameters (rebuilt when they move) — see engine.FlatParams."""

    _order: Sequence[str] = ()

    def _flat(self, device) -> FlatParams:
        f = getattr(self, "_flat_cache", None)
        own = dict(self.named_parameters())
        named = [(n, own[n]) for n in self._order]
        if f is not None and f.intact(named):
            f.sync_half()
            return f
        if torch.device(device).type != "cuda":
            raise B200Error(f"{type(self).__name__} runs on CUDA devices only (no CPU fallback)")
        object.__setattr__(self, "_flat_cache", FlatParams(named, device))
        return self._flat_cache

    def _drop_plan(self, device) -> Optional[DropPlan]:
        """The reference layers apply nn.Dropout to the attention probabilities and to both dense outputs
        (bert_model.py:338,373,451; probabilities handed in at ca_encoder.py:40-44 / ma_encoder.py:33-37): in train mode with
        p > 0 every forward draws one base seed (torch's CPU generator, so torch.manual_seed governs it) and the kernels derive
        the per-site masks from it; the backward regenerates them."""
        if not self.training or (self.p_hidden <= 0.0 and self.p_attn <= 0.0):
            return None
        return DropPlan(next_drop_seed(self, device), self.p_hidden, self.p_attn)

    def _need_grad(self, *inputs) -> bool:
        if not torch.is_grad_enabled():
            return False
        return any(p.requires_grad for p in self.parameters()) or any(t is not None and t.requires_grad for t in inputs)

    @staticmethod
    def _attn_views(f: FlatParams, prefix: str, kind: str, cross: bool) -> AttnWeights:
        n = [prefix + k for k in _ATTN]
        w = {"p": f.view16, "g": f.viewg}[kind]
        s = {"p": f.view32, "g": f.viewg}[kind]
        common = dict(wo=w(n[6]), bo=s(n[7]), g=s(n[8]), b=s(n[9]))
        if cross:
            return AttnWeights(wq=w(n[0]), bq=s(n[3]), wkv=w(n[1], (n[2],)), bkv=s(n[4], (n[5],)), **common)
        return AttnWeights(wqkv=w(n[0], (n[1], n[2])), bqkv=s(n[3], (n[4], n[5])), **common)

    @staticmethod
    def _ffn_views(f: FlatParams, kind: str) -> FfnWeights:
        w = {"p": f.view16, "g": f.viewg}[kind]
        s = {"p": f.view32, "g": f.viewg}[kind]
        return FfnWeights(w1=w(_FFN[0]), bf1=s(_FFN[1]), w2=w(_FFN[2]), bf2=s(_FFN[3]), g=s(_FFN[4]), b=s(_FFN[5]))


class _LayerFn(torch.autograd.Function):
    """One self-attention layer or one cross layer as a single autograd node."""

    @staticmethod
    def forward(ctx, layer, x, kv, self_bias, cross_bias, do_ffn, drop, need_grad, *params):
        # need_grad is decided by the module (grad MODE is not visible in ctx.needs_input_grad); drop: DropPlan or None
        f: FlatParams = layer._flat(x.device)
        B, N, H = x.shape
        heads, eps = layer.heads, layer.eps
        d = (lambda k: drop.layer(0, k)) if drop is not None else (lambda k: None)
        x32 = x.detach().to(F32).contiguous().view(B * N, H)
        x16 = _to16(x32)
        saved = []
        y16, y32, sv, _ = attn_block_fwd(layer._attn_views(f, "attention.", "p", False), x16, x32, B, N, heads, eps, self_bias, None,
                                         save=need_grad, drop_attn=d(DropPlan.ATTN), drop_hidden=d(DropPlan.ATTN_OUT))
        saved.append(sv)
        Nk = None
        if kv is not None:
            Nk = kv.shape[1]
            kv16 = _to16(kv.detach().to(F32).contiguous().view(B * Nk, kv.shape[2]))
            y16, y32, sv, _ = attn_block_fwd(layer._attn_views(f, "crossattention.", "p", True), y16, y32, B, N, heads, eps, cross_bias,
                                             None, save=need_grad, kv16=kv16, Sk=Nk, drop_attn=d(DropPlan.XATTN),
                                             drop_hidden=d(DropPlan.XATTN_OUT))
            saved.append(sv)
        if do_ffn:
            y16, y32, sv = ffn_block_fwd(layer._ffn_views(f, "p"), y16, y32, eps, save=need_grad, drop_hidden=d(DropPlan.FFN_OUT))
            saved.append(sv)
        ctx.layer, ctx.saved, ctx.dims = layer, (saved if need_grad else None), (B, N, Nk, H, kv.shape[2] if kv is not None else 0)
        ctx.biases, ctx.do_ffn, ctx.has_kv = (self_bias, cross_bias), do_ffn, kv is not None
        return y32.view(B, N, H)

    @staticmethod
    def backward(ctx, gy):
        layer, saved = ctx.layer, ctx.saved
        if saved is None:
            raise B200Error("backward through a forward that ran without grad")
        B, N, Nk, H, Hkv = ctx.dims
        f: FlatParams = layer._flat_cache
        dev = gy.device
        gy = gy.contiguous().to(F32)
        dy = torch.empty(B * N, H, dtype=F16, device=dev)
        scale = torch.empty(2, dtype=F32, device=dev)
        slot = torch.empty(1, dtype=torch.int32, device=dev)
        ops.scale_cast_grad(gy.view(-1), dy.view(-1), scale, slot, target=1024.0)
        inv = scale[1:2]
        keep, f.grad32 = f.grad32, torch.zeros_like(f.flat32)
        try:
            ws = ops.attn_bwd_workspace(B, layer.heads, N, dev)
            saved = list(saved)
            if ctx.do_ffn:
                dy = ffn_block_bwd(layer._ffn_views(f, "p"), layer._ffn_views(f, "g"), saved.pop(), dy, inv)
            dkv32 = None
            if ctx.has_kv:
                dy, dkv16 = attn_block_bwd(layer._attn_views(f, "crossattention.", "p", True), layer._attn_views(f, "crossattention.", "g", True),
                                           saved.pop(), dy, B, N, layer.heads, ctx.biases[1], None, inv, ws, Sk=Nk)
                dkv32 = ops.unscale_cast_grad(dkv16, torch.empty(B, Nk, Hkv, dtype=F32, device=dev), scale)
            dx16, _ = attn_block_bwd(layer._attn_views(f, "attention.", "p", False), layer._attn_views(f, "attention.", "g", False),
                                     saved.pop(), dy, B, N, layer.heads, ctx.biases[0], None, inv, ws)
            dx32 = ops.unscale_cast_grad(dx16, torch.empty(B, N, H, dtype=F32, device=dev), scale)
            grads = tuple(f.viewg(n) if f.params[n].requires_grad else None for n in f.names)
        finally:
            f.grad32 = keep
        ctx.saved = None
        return (None, dx32, dkv32, None, None, None, None, None) + grads


class BertSelfAttnLayer(_Packed):
    """= one HF BertLayer (bert_model.py:518-553).  forward(hidden_states, attention_mask, output_attentions=False) -> (y,)"""

    _order = ["attention." + k for k in _ATTN] + _FFN

    def __init__(self, bert_config, config=None):
        super().__init__()
        if bert_config.hidden_size != 64 * bert_config.num_attention_heads:
            raise B200Error("B200 cross-encoder layers require head_dim == 64")
        self.attention = BertAttention(bert_config)
        self.intermediate = BertIntermediate(bert_config)
        self.output = BertOutput(bert_config)
        self.heads, self.eps = bert_config.num_attention_heads, float(bert_config.layer_norm_eps)
        self.p_hidden = float(getattr(bert_config, "hidden_dropout_prob", 0.0) or 0.0)
        self.p_attn = float(getattr(bert_config, "attention_probs_dropout_prob", 0.0) or 0.0)

    def forward(self, hidden_states, attention_mask, output_attentions=False):
        if output_attentions:
            raise B200Error("output_attentions is not supported by the B200 cross-encoder layers")
        B, N, _ = hidden_states.shape
        bias = _mask_to_key_bias(attention_mask, B, N)
        own = dict(self.named_parameters())
        y = _LayerFn.apply(self, hidden_states, None, bias, None, True, self._drop_plan(hidden_states.device),
                           self._need_grad(hidden_states), *[own[n] for n in self._order])
        return (y,)


class BertCrossLayer(_Packed):
    """bert_model.py:456-515: self-attention block, cross-attention block (K/V projected from `encoder_hidden_states`
    of width ce_kv_hidden_size), FFN.  forward(hidden_states, encoder_hidden_states, attention_mask,
    encoder_attention_mask, output_attentions=False, do_ffn=True) -> (y,)"""

    _order = ["attention." + k for k in _ATTN] + ["crossattention." + k for k in _ATTN] + _FFN

    def __init__(self, config, ce_kv_hidden_size=None):
        super().__init__()
        if config.hidden_size != 64 * config.num_attention_heads:
            raise B200Error("B200 cross-encoder layers require head_dim == 64")
        self.attention = BertAttention(config)
        self.crossattention = BertAttention(config, kv_hidden_size=ce_kv_hidden_size)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)
        self.heads, self.eps = config.num_attention_heads, float(config.layer_norm_eps)
        self.p_hidden = float(getattr(config, "hidden_dropout_prob", 0.0) or 0.0)
        self.p_attn = float(getattr(config, "attention_probs_dropout_prob", 0.0) or 0.0)

    def forward(self, hidden_states, encoder_hidden_states, attention_mask, encoder_attention_mask, output_attentions=False,
                do_ffn=True):
        if output_attentions:
            raise B200Error("output_attentions is not supported by the B200 cross-encoder layers")
        B, N, _ = hidden_states.shape
        sb = _mask_to_key_bias(attention_mask, B, N)
        cb = _mask_to_key_bias(encoder_attention_mask, B, encoder_hidden_states.shape[1])
        own = dict(self.named_parameters())
        y = _LayerFn.apply(self, hidden_states, encoder_hidden_states, sb, cb, bool(do_ffn), self._drop_plan(hidden_states.device),
                           self._need_grad(hidden_states, encoder_hidden_states), *[own[n] for n in self._order])
        return (y,)


def _bert_config_from(config):
    from transformers import BertConfig
    return BertConfig(hidden_size=config.hidden_size, num_hidden_layers=config.num_cross_encoder_layers,
                      num_attention_heads=config.num_cross_encoder_heads, intermediate_size=config.intermediate_size,
                      max_position_embeddings=config.max_seq_length, hidden_dropout_prob=config.hidden_dropout_prob,
                      attention_probs_dropout_prob=config.attention_probs_dropout_prob)


def _init_weights(module):     # ma_encoder.py:14-23 / ca_encoder.py:17-26
    if isinstance(module, (nn.Linear, nn.Embedding)):
        module.weight.data.normal_(mean=0.0, std=0.02)
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)
    if isinstance(module, nn.Linear) and module.bias is not None:
        module.bias.data.zero_()


class MergeAttentionEncoder(nn.Module):
    """ma_encoder.py:10-71: concatenate the modalities along the sequence, run self-attention layers, split again."""

    def __init__(self, config):
        super().__init__()
        bc = _bert_config_from(config)
        self.cross_modal_layers = nn.ModuleList([BertSelfAttnLayer(bc, config) for _ in range(config.num_cross_encoder_layers)])
        self.cross_modal_layers.apply(_init_weights)

    def forward(self, attention_mask, text_feat=None, visual_feat=None, audio_feat=None):
        feats = [f for f in (text_feat, visual_feat, audio_feat) if f is not None]
        z = torch.cat(feats, dim=1)
        cat_mask = torch.cat([attention_mask] * len(feats), dim=1)
        ext = (1.0 - cat_mask[:, None, None, :].to(torch.float)) * -1000000.0
        for layer in self.cross_modal_layers:
            z = layer(z, ext)[0]
        outs = list(torch.chunk(z, len(feats), dim=1))
        res = [outs.pop(0) if f is not None else None for f in (text_feat, visual_feat, audio_feat)]
        return tuple(res)


class CoAttentionEncoder(nn.Module):
    """ca_encoder.py:13-77: one BertCrossLayer stack per modality; K/V come from the other modalities (concatenated on the
    hidden dim when all three are present)."""

    def __init__(self, config):
        super().__init__()
        bc = _bert_config_from(config)
        mk = lambda: nn.ModuleList([BertCrossLayer(bc, ce_kv_hidden_size=config.ce_kv_hidden_size)
                                    for _ in range(config.num_cross_encoder_layers)])
        self.cross_modal_visual_layers = mk()
        self.cross_modal_visual_layers.apply(_init_weights)
        self.cross_modal_text_layers = mk()
        self.cross_modal_text_layers.apply(_init_weights)
        self.cross_modal_audio_layers = mk()
        self.cross_modal_audio_layers.apply(_init_weights)

    def forward(self, attention_mask, text_feat=None, visual_feat=None, audio_feat=None):
        ext = (1.0 - attention_mask[:, None, None, :].to(torch.float)) * -1000000.0
        t, v, a = text_feat, visual_feat, audio_feat
        for tl, vl, al in zip(self.cross_modal_text_layers, self.cross_modal_visual_layers, self.cross_modal_audio_layers):
            if t is None:
                v, a = vl(v, a, ext, ext)[0], al(a, v, ext, ext)[0]
            elif v is None:
                t, a = tl(t, a, ext, ext)[0], al(a, t, ext, ext)[0]
            elif a is None:
                t, v = tl(t, v, ext, ext)[0], vl(v, t, ext, ext)[0]
            else:
                av, at, tv = torch.cat((a, v), dim=-1), torch.cat((a, t), dim=-1), torch.cat((t, v), dim=-1)
                t, v, a = tl(t, av, ext, ext)[0], vl(v, at, ext, ext)[0], al(a, tv, ext, ext)[0]
        return t, v, a


class _ProjFn(torch.autograd.Function):
    """y = LayerNorm(x W^T + b)  (linear_projector.py:21-30; dropout is applied by the caller module)."""

    @staticmethod
    def forward(ctx, x, W, b, gamma, beta, eps, bf16=False):
        """`bf16`: the forward contraction runs on bf16 operands (BASELINE config 4's bf16 arm — what a bf16 autocast would feed
        the tensor cores; same rate, 8 mantissa bits); the backward keeps fp16 operands under the loss scale."""
        shp = x.shape
        x32 = x.detach().to(F32).contiguous().view(-1, shp[-1])
        M, dev = x32.shape[0], x.device
        x16, W16 = _to16(x32), _to16(W.detach().contiguous())
        H = W.shape[0]
        pre = torch.empty(M, H, dtype=F32, device=dev)
        if bf16:
            ops.gemm_bf16(ops.cast_f32_to_bf16(x32), ops.cast_f32_to_bf16(W.detach().to(F32).contiguous()), pre, epilogue=ops.EPI_BIAS,
                          bias=b.detach())
        else:
            ops.gemm(x16, W16, pre, epilogue=ops.EPI_BIAS, bias=b.detach())
        mean, rstd = torch.empty(M, dtype=F32, device=dev), torch.empty(M, dtype=F32, device=dev)
        y32 = torch.empty(M, H, dtype=F32, device=dev)
        ops.layernorm_fwd(pre, gamma.detach(), beta.detach(), eps, y32=y32, mean=mean, rstd=rstd)
        ctx.save_for_backward(x16, W16, pre, mean, rstd, gamma.detach())
        ctx.shp = shp
        return y32.view(*shp[:-1], H)

    @staticmethod
    def backward(ctx, gy):
        x16, W16, pre, mean, rstd, gamma = ctx.saved_tensors
        M, H = pre.shape
        K, dev = x16.shape[1], gy.device
        dy = torch.empty(M, H, dtype=F16, device=dev)
        scale = torch.empty(2, dtype=F32, device=dev)
        slot = torch.empty(1, dtype=torch.int32, device=dev)
        ops.scale_cast_grad(gy.contiguous().to(F32).view(-1), dy.view(-1), scale, slot, target=1024.0)
        inv = scale[1:2]
        d_pre = torch.empty(M, H, dtype=F16, device=dev)
        dg, db, dbias = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
        ops.layernorm_bwd(dy, pre, mean, rstd, gamma, d_pre, dg, db, dbias=dbias, alpha=inv)
        dW = torch.zeros(H, K, dtype=F32, device=dev)
        ops.gemm(d_pre, x16, dW, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv, k_splits=ops.wgrad_splits(H, K, M))
        dx16 = torch.empty(M, K, dtype=F16, device=dev)
        ops.gemm(d_pre, W16, dx16, b_layout=1)
        dx = ops.unscale_cast_grad(dx16, torch.empty(ctx.shp, dtype=F32, device=dev), scale)
        return dx, dW, dbias, dg, db, None, None


class LinearProjector(nn.Module):
    """linear_projector.py:5-30 — Dropout(LayerNorm(Linear(x))) per modality."""

    def __init__(self, config):
        super().__init__()
        # "bf16": run the projection GEMMs on bf16 operands (config attribute `b200_operand_dtype`, default "fp16")
        self.operand_dtype = getattr(config, "b200_operand_dtype", "fp16")
        if self.operand_dtype not in ("fp16", "bf16"):
            raise B200Error(f"b200_operand_dtype must be 'fp16' or 'bf16', got {self.operand_dtype!r}")
        for name, dim in (("text", config.hidden_size), ("vis", config.hidden_size_vis), ("audio", config.hidden_size_audio)):
            if dim % 8:
                raise B200Error(f"LinearProjector input width {dim} must be a multiple of 8")
            setattr(self, f"proj_{name}", nn.Linear(dim, config.hidden_size))
            setattr(self, f"layernorm_{name}", nn.LayerNorm(config.hidden_size))
            setattr(self, f"dropout_{name}", nn.Dropout(config.hidden_dropout_prob))

    def _one(self, name, x):
        if x is None:
            return None
        lin, ln = getattr(self, f"proj_{name}"), getattr(self, f"layernorm_{name}")
        return getattr(self, f"dropout_{name}")(_ProjFn.apply(x, lin.weight, lin.bias, ln.weight, ln.bias, float(ln.eps),
                                                              self.operand_dtype == "bf16"))

    def forward(self, text_feature=None, vis_feature=None, audio_feature=None):
        return self._one("text", text_feature), self._one("vis", vis_feature), self._one("audio", audio_feature)
