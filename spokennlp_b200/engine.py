"""Host-side orchestration of the encoder hot path: parameter packing and the per-layer
kernel schedule (forward and backward) over the C ABI.  No arithmetic happens in Python.

Data layout in HBM:
  * one flat fp32 buffer holds every parameter (HF order below); the module's nn.Parameters are VIEWS into it, so
    checkpoints load/save with HF keys while query/key/value sit adjacent and form the packed [3H,H] QKV weight;
  * one flat fp16 buffer mirrors it (tensor-core operands); one flat fp32 buffer holds the gradients;
  * activations are token-major [B*S, width] fp16; Q/K/V live packed in one [B*S, 3H] buffer and are addressed by
    TMA coordinates (no head split / permute); pre-LayerNorm sums are kept in fp32.

Reference schedule being replaced: BertLayer.forward (mmvts/src/models/cross_encoder/bert_model.py:518-553) and its
autograd; embeddings (:184-210).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops

Tensor = torch.Tensor
ALIGN = 64  # elements


def layer_param_names(i: int) -> List[str]:
    p = f"encoder.layer.{i}."
    return [p + "attention.self.query.weight", p + "attention.self.key.weight", p + "attention.self.value.weight",
            p + "attention.self.query.bias", p + "attention.self.key.bias", p + "attention.self.value.bias",
            p + "attention.output.dense.weight", p + "attention.output.dense.bias",
            p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias",
            p + "intermediate.dense.weight", p + "intermediate.dense.bias",
            p + "output.dense.weight", p + "output.dense.bias",
            p + "output.LayerNorm.weight", p + "output.LayerNorm.bias"]


EMB_NAMES = ["embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight",
             "embeddings.token_type_embeddings.weight", "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias"]


class FlatParams:
    """Packs named fp32 parameters, in the given order, into one flat buffer (+ fp16 mirror, + optional grad buffer)
    and re-points the nn.Parameters at views of it."""

    def __init__(self, named: Sequence[Tuple[str, torch.nn.Parameter]], device):
        self.names = [n for n, _ in named]
        self.params = {n: p for n, p in named}
        self.offsets: Dict[str, int] = {}
        off = 0
        for n in self.names:                      # every offset is 64-element aligned; tensors whose size is a
            self.offsets[n] = off                 # multiple of 64 (all H x H / H-sized ones) therefore stay adjacent
            off = (off + self.params[n].numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.flat32 = torch.zeros(off, dtype=torch.float32, device=device)
        self.flat16 = torch.empty(off, dtype=torch.float16, device=device)
        self.grad32: Optional[Tensor] = None
        with torch.no_grad():
            for n, p in self.params.items():
                v = self.view32(n)
                v.copy_(p.data.to(device=device, dtype=torch.float32))
                p.data = v
        self.signature = self._sig()
        self.version = -1
        self.sync_half()

    def _sig(self):
        return tuple(p.data_ptr() for p in self.params.values())

    def intact(self) -> bool:
        return self._sig() == self.signature

    def _view(self, flat: Tensor, name: str, n_extra: Sequence[str] = ()) -> Tensor:
        p = self.params[name]
        o = self.offsets[name]
        if not n_extra:
            return flat[o:o + p.numel()].view(p.shape)
        rows = p.shape[0] + sum(self.params[e].shape[0] for e in n_extra)
        numel = p.numel() + sum(self.params[e].numel() for e in n_extra)
        end = o
        for e in (name,) + tuple(n_extra):        # the packed view is only valid if the members are adjacent
            assert self.offsets[e] == end, f"{e} is not adjacent in the flat buffer"
            end += self.params[e].numel()
        return flat[o:o + numel].view(rows, *p.shape[1:])

    def view32(self, name, extra=()):
        return self._view(self.flat32, name, extra)

    def view16(self, name, extra=()):
        return self._view(self.flat16, name, extra)

    def viewg(self, name, extra=()):
        return self._view(self.grad32, name, extra)

    def cur_version(self) -> int:
        return sum(p._version for p in self.params.values())

    def sync_half(self, force: bool = False) -> None:
        """Refresh the fp16 mirror if any parameter was modified in place since the last refresh."""
        v = self.cur_version()
        if force or v != self.version:
            ops.cast_f32_to_f16(self.flat32, self.flat16)
            self.version = v

    def ensure_grad(self, zero: bool = True) -> Tensor:
        if self.grad32 is None:
            self.grad32 = torch.zeros_like(self.flat32)
        elif zero:
            self.grad32.zero_()
        return self.grad32


@dataclass
class LayerViews:
    wqkv: Tensor; bqkv: Tensor; wo: Tensor; bo: Tensor; g1: Tensor; b1: Tensor
    w1: Tensor; bf1: Tensor; w2: Tensor; bf2: Tensor; g2: Tensor; b2: Tensor


@dataclass
class LayerSaved:
    x_in: Tensor = None; qkv: Tensor = None; ctx: Tensor = None; lse2: Tensor = None
    pre1: Tensor = None; mean1: Tensor = None; rstd1: Tensor = None; ln1: Tensor = None
    z: Tensor = None; h: Tensor = None          # z holds gelu'(pre-activation), h = gelu(pre-activation)
    pre2: Tensor = None; mean2: Tensor = None; rstd2: Tensor = None


@dataclass
class Saved:
    B: int = 0; S: int = 0
    ids: Optional[Tensor] = None; tt: Optional[Tensor] = None; pos: Optional[Tensor] = None
    key_bias: Optional[Tensor] = None; kv_len: Optional[Tensor] = None
    layers: List[LayerSaved] = field(default_factory=list)
    out: Tensor = None


class EncoderEngine:
    """Runs embeddings + L encoder layers on the packed parameters."""

    def __init__(self, flat: FlatParams, hidden: int, heads: int, inter: int, n_layers: int, eps: float,
                 prefix_layers: str = "encoder.layer."):
        assert hidden == heads * 64, "the sm_100a attention kernels are specialised for head_dim 64"
        self.flat, self.H, self.heads, self.I, self.L, self.eps = flat, hidden, heads, inter, n_layers, eps
        self.prefix = prefix_layers

    # ---- views -------------------------------------------------------------------------------------------------
    def _lv(self, i: int, kind: str) -> LayerViews:
        n = layer_param_names(i)
        f = self.flat
        w = {"p": f.view16, "g": f.viewg}[kind]
        s = {"p": f.view32, "g": f.viewg}[kind]
        return LayerViews(wqkv=w(n[0], (n[1], n[2])), bqkv=s(n[3], (n[4], n[5])), wo=w(n[6]), bo=s(n[7]), g1=s(n[8]),
                          b1=s(n[9]), w1=w(n[10]), bf1=s(n[11]), w2=w(n[12]), bf2=s(n[13]), g2=s(n[14]), b2=s(n[15]))

    def layer(self, i):
        return self._lv(i, "p")

    def layer_grads(self, i):
        return self._lv(i, "g")

    # ---- forward -----------------------------------------------------------------------------------------------
    def embed(self, ids, tt, pos, inputs_embeds, B, S):
        """Returns the embedding output twice: fp16 (tensor-core operand) and fp32 (residual stream)."""
        f = self.flat
        y32 = torch.empty(B * S, self.H, dtype=torch.float32, device=f.flat32.device)
        y16 = ops.embed_ln_fwd(ids, tt, pos, inputs_embeds, f.view32(EMB_NAMES[0]), f.view32(EMB_NAMES[1]),
                               f.view32(EMB_NAMES[2]), f.view32(EMB_NAMES[3]), f.view32(EMB_NAMES[4]), self.eps, B * S, S,
                               self.H, y32=y32)
        return y16, y32

    def layer_forward(self, p: LayerViews, x: Tensor, x32: Tensor, B: int, S: int, key_bias, kv_len, save: bool,
                      want_probs: bool = False):
        """x: fp16 layer input (GEMM operand), x32: the same activations in fp32 (residual stream: keeping the skip
        connection un-rounded is what holds the 12-layer hidden-state error under 1e-3)."""
        H, I, M, dev = self.H, self.I, B * S, x.device
        f16, f32 = torch.float16, torch.float32
        sv = LayerSaved() if save else None
        qkv = torch.empty(M, 3 * H, dtype=f16, device=dev)
        ops.gemm(x, p.wqkv, qkv, epilogue=ops.EPI_BIAS, bias=p.bqkv)
        ctx = torch.empty(M, H, dtype=f16, device=dev)
        lse2 = torch.empty(B, self.heads, S, dtype=f32, device=dev) if (save or want_probs) else None
        ops.attn_fwd(qkv, qkv, ctx, B, self.heads, S, S, q_col0=0, k_col0=H, v_col0=2 * H, key_bias=key_bias, kv_len=kv_len,
                     lse2=lse2)
        probs = ops.attn_probs(qkv, qkv, lse2, B, self.heads, S, S, q_col0=0, k_col0=H, key_bias=key_bias) if want_probs else None
        pre1 = torch.empty(M, H, dtype=f32, device=dev)
        ops.gemm(ctx, p.wo, pre1, epilogue=ops.EPI_BIAS_RES32, bias=p.bo, aux=x32)
        mean1 = torch.empty(M, dtype=f32, device=dev) if save else None
        rstd1 = torch.empty(M, dtype=f32, device=dev) if save else None
        ln1_32 = torch.empty(M, H, dtype=f32, device=dev)
        ln1 = ops.layernorm_fwd(pre1, p.g1, p.b1, self.eps, y32=ln1_32, mean=mean1, rstd=rstd1)
        h = torch.empty(M, I, dtype=f16, device=dev)
        z = torch.empty(M, I, dtype=f16, device=dev) if save else None
        ops.gemm(ln1, p.w1, h, epilogue=ops.EPI_BIAS_GELU, bias=p.bf1, out2=z)
        pre2 = torch.empty(M, H, dtype=f32, device=dev)
        ops.gemm(h, p.w2, pre2, epilogue=ops.EPI_BIAS_RES32, bias=p.bf2, aux=ln1_32)
        mean2 = torch.empty(M, dtype=f32, device=dev) if save else None
        rstd2 = torch.empty(M, dtype=f32, device=dev) if save else None
        out32 = torch.empty(M, H, dtype=f32, device=dev)
        out = ops.layernorm_fwd(pre2, p.g2, p.b2, self.eps, y32=out32, mean=mean2, rstd=rstd2)
        if save:
            sv.x_in, sv.qkv, sv.ctx, sv.lse2 = x, qkv, ctx, lse2
            sv.pre1, sv.mean1, sv.rstd1, sv.ln1 = pre1, mean1, rstd1, ln1
            sv.z, sv.h, sv.pre2, sv.mean2, sv.rstd2 = z, h, pre2, mean2, rstd2
        return out, out32, sv, probs

    def forward(self, ids, tt, pos, inputs_embeds, key_bias, kv_len, B: int, S: int, *, save: bool,
                want_hidden: bool = False, want_probs: bool = False):
        self.flat.sync_half()
        x, x32 = self.embed(ids, tt, pos, inputs_embeds, B, S)
        saved = Saved(B=B, S=S, ids=ids, tt=tt, pos=pos, key_bias=key_bias, kv_len=kv_len) if save else None
        hiddens, probs_all = ([x32] if want_hidden else None), ([] if want_probs else None)
        for i in range(self.L):
            x, x32, sv, probs = self.layer_forward(self.layer(i), x, x32, B, S, key_bias, kv_len, save, want_probs)
            if save:
                saved.layers.append(sv)
            if want_hidden:
                hiddens.append(x32)
            if want_probs:
                probs_all.append(probs)
        if save:
            saved.out = x
        return x, x32, saved, hiddens, probs_all

    # ---- backward ----------------------------------------------------------------------------------------------
    def layer_backward(self, p: LayerViews, g: LayerViews, sv: LayerSaved, dy: Tensor, B: int, S: int, key_bias, kv_len,
                       inv_scale: Optional[Tensor], ws: Tensor) -> Tensor:
        """dy: fp16 gradient wrt the layer output (scaled by the loss scale).  Weight/bias/LN gradients are accumulated
        (+=) into the fp32 views `g`, multiplied by *inv_scale.  Returns the gradient wrt the layer input."""
        H, I, M, dev = self.H, self.I, B * S, dy.device
        f16 = torch.float16
        # output LayerNorm  <- BertOutput (bert_model.py:449-453)
        d_pre2 = torch.empty(M, H, dtype=f16, device=dev)
        ops.layernorm_bwd(dy, sv.pre2, sv.mean2, sv.rstd2, p.g2, d_pre2, g.g2, g.b2, dbias=g.bf2, alpha=inv_scale)
        ops.gemm(d_pre2, sv.h, g.w2, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv_scale,
                 k_splits=ops.wgrad_splits(H, I, M))
        dz = torch.empty(M, I, dtype=f16, device=dev)
        ops.gemm(d_pre2, p.w2, dz, b_layout=1, epilogue=ops.EPI_DGELU, aux=sv.z)
        ops.colsum(dz, g.bf1, inv_scale)
        ops.gemm(dz, sv.ln1, g.w1, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv_scale,
                 k_splits=ops.wgrad_splits(I, H, M))
        d_ln1 = torch.empty(M, H, dtype=f16, device=dev)
        ops.gemm(dz, p.w1, d_ln1, b_layout=1, epilogue=ops.EPI_ADD, aux=d_pre2)
        # attention output LayerNorm  <- BertSelfOutput (:371-375)
        d_pre1 = torch.empty(M, H, dtype=f16, device=dev)
        ops.layernorm_bwd(d_ln1, sv.pre1, sv.mean1, sv.rstd1, p.g1, d_pre1, g.g1, g.b1, dbias=g.bo, alpha=inv_scale)
        ops.gemm(d_pre1, sv.ctx, g.wo, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv_scale,
                 k_splits=ops.wgrad_splits(H, H, M))
        dctx = torch.empty(M, H, dtype=f16, device=dev)
        ops.gemm(d_pre1, p.wo, dctx, b_layout=1)
        # attention core (:309-350)
        dqkv = torch.empty(M, 3 * H, dtype=f16, device=dev)
        ops.attn_bwd(sv.qkv, sv.qkv, dctx, sv.ctx, sv.lse2, dqkv, dqkv, ws, B, self.heads, S, S, q_col0=0, k_col0=H,
                     v_col0=2 * H, dq_col0=0, dk_col0=H, dv_col0=2 * H, key_bias=key_bias, kv_len=kv_len)
        ops.colsum(dqkv, g.bqkv, inv_scale)
        ops.gemm(dqkv, sv.x_in, g.wqkv, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv_scale,
                 k_splits=ops.wgrad_splits(3 * H, H, M))
        dx = torch.empty(M, H, dtype=f16, device=dev)
        ops.gemm(dqkv, p.wqkv, dx, b_layout=1, epilogue=ops.EPI_ADD, aux=d_pre1)
        return dx

    def backward(self, saved: Saved, dy: Tensor, inv_scale: Optional[Tensor], *, embeddings: bool = True,
                 after_layer=None) -> None:
        """Accumulates all parameter gradients into flat.grad32 (must exist).  `after_layer(i)` is called once layer
        i's gradients are complete (the data-parallel trainer launches that layer's allreduce there)."""
        B, S = saved.B, saved.S
        f = self.flat
        ws = ops.attn_bwd_workspace(B, self.heads, S, dy.device)
        for i in reversed(range(self.L)):
            dy = self.layer_backward(self.layer(i), self.layer_grads(i), saved.layers[i], dy, B, S, saved.key_bias,
                                     saved.kv_len, inv_scale, ws)
            saved.layers[i] = None          # release this layer's activations
            if after_layer is not None:
                after_layer(i)
        if embeddings and saved.ids is not None:
            ops.embed_ln_bwd(dy, None, saved.ids, saved.tt, saved.pos, f.view32(EMB_NAMES[0]), f.view32(EMB_NAMES[1]),
                             f.view32(EMB_NAMES[2]), f.view32(EMB_NAMES[3]), f.viewg(EMB_NAMES[0]), f.viewg(EMB_NAMES[1]),
                             f.viewg(EMB_NAMES[2]), f.viewg(EMB_NAMES[3]), f.viewg(EMB_NAMES[4]), inv_scale, self.eps, B * S,
                             S, self.H)
        if after_layer is not None:
            after_layer(-1)
