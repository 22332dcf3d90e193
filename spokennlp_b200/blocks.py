"""Kernel schedules of the two building blocks every encoder layer of the path is made of — an attention block
(projections -> fused attention -> output dense + residual + LayerNorm) and a feed-forward block (dense + GELU ->
dense + residual + LayerNorm) — forward and backward, over explicit weight views.  `BertLayer` /
`BertSelfAttnLayer` = attention + FFN; `BertCrossLayer` = self-attention + cross-attention + FFN
(mmvts/src/models/cross_encoder/bert_model.py:456-553).  No arithmetic happens here: every line enqueues a kernel.

Conventions: `x16` is the fp16 tensor-core operand copy of the block input, `x32` the same activations in fp32 (the
residual stream); gradients flowing between blocks are fp16, scaled by the loss scale; parameter gradients are
accumulated (+=) into fp32 views and multiplied by `*inv_scale` inside the producing kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import ops

Tensor = torch.Tensor
F16, F32 = torch.float16, torch.float32


class Experimental:
    """Host-side schedule switches: which fused entry points a layer calls.  Nothing here is library state — each switch picks
    between two sequences of C-ABI calls that compute the same thing.  All four were validated and measured on a B200 in
    round 2 (profiles/r02a_variants_ab.jsonl, bench_r02a_*.json, r02c_*, bench_r02i_*) and are ON by default:
      resadd : output-dense + residual through b200_gemm_f16_resadd (in place on the fp32 residual stream, no aux reads):
               out-proj 51 -> 28 us, FFN-down 81 -> 74 us per layer at the bench shape (64 -> 29 / 84 -> 74 with dropout)
      delta  : attention-backward row statistic fused into the output projection's dgrad (b200_gemm_f16_dgrad_delta): -12 us/layer
      colsum : bias gradient of the FFN-up dense summed inside the FFN-down dgrad's dGELU epilogue (b200_gemm_f16_dgelu_colsum)
      dq16   : attention backward accumulates dQ as fp16 TMA reduce-adds in place (B200_ATTN_BWD_DQ_HALF): no fp32 accumulator,
               memset or cast pass, at most Sk/128 roundings per element instead of one; 13.99 -> 13.70 ms/step on one box
    `B200_EXP` (comma-separated) names the switches to turn on instead of the default set; `B200_EXP=none` is the round-1
    schedule.  (Measured and dropped in round 2: a stream-K schedule for the resadd GEMMs, lane-elected mbarrier WAITS in the
    attention kernels, and sixteen-warp attention kernels — DESIGN.md §9.)"""
    DEFAULT = "resadd,delta,colsum,dq16"
    NAMES = ("resadd", "delta", "colsum", "dq16")
    resadd = delta = colsum = dq16 = False

    @classmethod
    def from_env(cls, value: Optional[str] = None) -> None:
        import os
        raw = os.environ.get("B200_EXP") if value is None else value
        if raw is None or raw.strip() == "default":
            raw = cls.DEFAULT
        names = {n.strip() for n in raw.split(",") if n.strip() and n.strip() != "none"}
        unknown = names - set(cls.NAMES)
        if unknown:
            raise ValueError(f"B200_EXP: unknown switch(es) {sorted(unknown)}")
        for n in cls.NAMES:
            setattr(cls, n, n in names)

    @classmethod
    def active(cls):
        return [n for n in cls.NAMES if getattr(cls, n)]


Experimental.from_env()


def _out_dense_residual(h16: Tensor, w: Tensor, bias: Tensor, x32: Tensor, drop, owns_residual: bool) -> Tensor:
    """pre = dropout(h16 @ w^T + bias) + x32, fp32 (BertSelfOutput / BertOutput before their LayerNorm).  `owns_residual`:
    nobody else reads x32 afterwards, so the in-place variant may accumulate into it."""
    if Experimental.resadd and owns_residual:
        return ops.gemm_resadd(h16, w, x32, bias, drop=drop)
    pre = torch.empty(h16.shape[0], w.shape[0], dtype=F32, device=h16.device)
    ops.gemm(h16, w, pre, epilogue=ops.EPI_BIAS_RES32, bias=bias, aux=x32, drop=drop)
    return pre


@dataclass
class AttnWeights:
    """Self-attention: wqkv [3H,H] / bqkv packed.  Cross-attention: wq [H,H], bq and wkv [2H,Hkv], bkv (key and value
    adjacent).  wo/bo: output dense; g/b: LayerNorm."""
    wo: Tensor
    bo: Tensor
    g: Tensor
    b: Tensor
    wqkv: Optional[Tensor] = None
    bqkv: Optional[Tensor] = None
    wq: Optional[Tensor] = None
    bq: Optional[Tensor] = None
    wkv: Optional[Tensor] = None
    bkv: Optional[Tensor] = None


@dataclass
class FfnWeights:
    w1: Tensor
    bf1: Tensor
    w2: Tensor
    bf2: Tensor
    g: Tensor
    b: Tensor


@dataclass
class AttnSaved:
    x16: Tensor = None
    kv16: Tensor = None          # cross-attention only: the fp16 K/V source
    q: Tensor = None             # [Mq, 3H] packed (self) or [Mq, H] (cross)
    kv: Tensor = None            # cross-attention only: [Mk, 2H]
    ctx: Tensor = None
    lse2: Tensor = None
    pre: Tensor = None
    mean: Tensor = None
    rstd: Tensor = None
    pack: Optional[ops.RowIndex] = None          # packed rows
    drop_attn: Optional[ops.Dropout] = None      # dropout on the attention probabilities (bert_model.py:338)
    drop_hidden: Optional[ops.Dropout] = None    # dropout on the output dense, before the residual (bert_model.py:373)


@dataclass
class FfnSaved:
    x16: Tensor = None
    dact: Tensor = None          # gelu'(pre-activation)
    h: Tensor = None
    pre: Tensor = None
    mean: Tensor = None
    rstd: Tensor = None
    drop_hidden: Optional[ops.Dropout] = None    # bert_model.py:451


def attn_block_fwd(p: AttnWeights, x16: Tensor, x32: Tensor, B: int, Sq: int, heads: int, eps: float, key_bias, kv_len, *,
                   save: bool, want_probs: bool = False, kv16: Optional[Tensor] = None, Sk: Optional[int] = None,
                   drop_attn: Optional[ops.Dropout] = None, drop_hidden: Optional[ops.Dropout] = None, owns_residual: bool = False,
                   pack: Optional[ops.RowIndex] = None):
    """bert_model.py:259-375 (BertSelfAttention + BertSelfOutput).  Returns (y16, y32, saved, probs).  `probs` are the
    probabilities BEFORE dropout.  `owns_residual`: the caller hands x32 over (it is not a hidden state anybody keeps).
    `pack`: the rows are PACKED valid tokens (Sq = the padded length)."""
    H, Mq, dev = heads * 64, x16.shape[0], x16.device
    cross = kv16 is not None
    Sk = Sk if cross else Sq
    if cross:
        q = torch.empty(Mq, H, dtype=F16, device=dev)
        ops.gemm(x16, p.wq, q, epilogue=ops.EPI_BIAS, bias=p.bq)
        kv = torch.empty(B * Sk, 2 * H, dtype=F16, device=dev)
        ops.gemm(kv16, p.wkv, kv, epilogue=ops.EPI_BIAS, bias=p.bkv)
        cols = dict(q_col0=0, k_col0=0, v_col0=H)
    else:
        q = torch.empty(Mq, 3 * H, dtype=F16, device=dev)
        ops.gemm(x16, p.wqkv, q, epilogue=ops.EPI_BIAS, bias=p.bqkv)
        kv = q
        cols = dict(q_col0=0, k_col0=H, v_col0=2 * H)
    ctx = torch.empty(Mq, H, dtype=F16, device=dev)
    lse2 = torch.empty(B, heads, Sq, dtype=F32, device=dev) if (save or want_probs) else None
    ops.attn_fwd(q, kv, ctx, B, heads, Sq, Sk, key_bias=key_bias, kv_len=kv_len, lse2=lse2, drop=drop_attn, pack=pack, **cols)
    probs = None
    if want_probs:
        if pack is not None:
            raise ops.L.B200Error("output_attentions is not available on packed rows")
        probs = ops.attn_probs(q, kv, lse2, B, heads, Sq, Sk, q_col0=cols["q_col0"], k_col0=cols["k_col0"], key_bias=key_bias)
    pre = _out_dense_residual(ctx, p.wo, p.bo, x32, drop_hidden, owns_residual)
    mean = torch.empty(Mq, dtype=F32, device=dev) if save else None
    rstd = torch.empty(Mq, dtype=F32, device=dev) if save else None
    y32 = torch.empty(Mq, H, dtype=F32, device=dev)
    y16 = ops.layernorm_fwd(pre, p.g, p.b, eps, y32=y32, mean=mean, rstd=rstd)
    sv = AttnSaved(x16=x16, kv16=kv16, q=q, kv=kv if cross else None, ctx=ctx, lse2=lse2, pre=pre, mean=mean, rstd=rstd,
                   drop_attn=drop_attn, drop_hidden=drop_hidden, pack=pack) if save else None
    return y16, y32, sv, probs


def ffn_block_fwd(p: FfnWeights, x16: Tensor, x32: Tensor, eps: float, *, save: bool, drop_hidden: Optional[ops.Dropout] = None,
                  owns_residual: bool = False):
    """bert_model.py:436-453 (BertIntermediate + BertOutput).  Returns (y16, y32, saved)."""
    M, H, dev = x16.shape[0], x16.shape[1], x16.device
    inter = p.w1.shape[0]
    h = torch.empty(M, inter, dtype=F16, device=dev)
    dact = torch.empty(M, inter, dtype=F16, device=dev) if save else None
    ops.gemm(x16, p.w1, h, epilogue=ops.EPI_BIAS_GELU, bias=p.bf1, out2=dact)
    pre = _out_dense_residual(h, p.w2, p.bf2, x32, drop_hidden, owns_residual)
    mean = torch.empty(M, dtype=F32, device=dev) if save else None
    rstd = torch.empty(M, dtype=F32, device=dev) if save else None
    y32 = torch.empty(M, H, dtype=F32, device=dev)
    y16 = ops.layernorm_fwd(pre, p.g, p.b, eps, y32=y32, mean=mean, rstd=rstd)
    sv = FfnSaved(x16=x16, dact=dact, h=h, pre=pre, mean=mean, rstd=rstd, drop_hidden=drop_hidden) if save else None
    return y16, y32, sv


def _ln_bwd(dy, dy2, sv, gamma, dgamma, dbeta, dbias, inv_scale):
    """LayerNorm backward of a block's closing `LayerNorm(dropout(dense(.)) + x)`.  Returns (d_pre, d_dense): the gradient
    wrt the pre-LayerNorm sum (what the residual path carries) and wrt the dense output (the same tensor unless hidden
    dropout is on, in which case it is d_pre times the regenerated forward mask)."""
    d_pre = torch.empty(dy.shape, dtype=F16, device=dy.device)
    dr = sv.drop_hidden
    if dr is None or dr.p <= 0.0:
        ops.layernorm_bwd(dy, sv.pre, sv.mean, sv.rstd, gamma, d_pre, dgamma, dbeta, dy2=dy2, dbias=dbias, alpha=inv_scale)
        return d_pre, d_pre
    d_den = torch.empty_like(d_pre)
    ops.layernorm_bwd(dy, sv.pre, sv.mean, sv.rstd, gamma, d_pre, dgamma, dbeta, dy2=dy2, dbias=dbias, alpha=inv_scale, dx_drop=d_den,
                      drop=dr)
    return d_pre, d_den


def ffn_block_bwd(p: FfnWeights, g: FfnWeights, sv: FfnSaved, dy: Tensor, inv_scale, dy2: Optional[Tensor] = None) -> Tensor:
    """Backward of ffn_block_fwd.  dy (+dy2): gradient wrt the block output; returns the gradient wrt the block input
    (dense path + residual path)."""
    M, H = dy.shape
    inter, dev = p.w1.shape[0], dy.device
    d_pre, d_den = _ln_bwd(dy, dy2, sv, p.g, g.g, g.b, g.bf2, inv_scale)
    ops.gemm(d_den, sv.h, g.w2, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv_scale, k_splits=ops.wgrad_splits(H, inter, M))
    dz = torch.empty(M, inter, dtype=F16, device=dev)
    if Experimental.colsum:                      # bias gradient of the intermediate dense taken in the dgrad's epilogue
        ops.gemm_dgelu_colsum(d_den, p.w2, sv.dact, dz, g.bf1, inv_scale)
    else:
        ops.gemm(d_den, p.w2, dz, b_layout=1, epilogue=ops.EPI_DGELU, aux=sv.dact)
        ops.colsum(dz, g.bf1, inv_scale)
    ops.gemm(dz, sv.x16, g.w1, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv_scale, k_splits=ops.wgrad_splits(inter, H, M))
    dx = torch.empty(M, H, dtype=F16, device=dev)
    ops.gemm(dz, p.w1, dx, b_layout=1, epilogue=ops.EPI_ADD, aux=d_pre)
    return dx


def attn_block_bwd(p: AttnWeights, g: AttnWeights, sv: AttnSaved, dy: Tensor, B: int, Sq: int, heads: int, key_bias, kv_len,
                   inv_scale, ws: Tensor, *, Sk: Optional[int] = None, dy2: Optional[Tensor] = None):
    """Backward of attn_block_fwd.  Returns (dx, dkv_src): gradient wrt the block input (projection path + residual
    path) and, for cross-attention, wrt the fp16 K/V source (else None)."""
    H, Mq, dev = heads * 64, dy.shape[0], dy.device
    cross = sv.kv16 is not None
    Sk = Sk if cross else Sq
    d_pre, d_den = _ln_bwd(dy, dy2, sv, p.g, g.g, g.b, g.bo, inv_scale)
    ops.gemm(d_den, sv.ctx, g.wo, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv_scale, k_splits=ops.wgrad_splits(H, H, Mq))
    dctx = torch.empty(Mq, H, dtype=F16, device=dev)
    fused_delta = Experimental.delta and not cross and sv.pack is None     # (the fused epilogue maps rows to sequences by division)
    if fused_delta:
        ops.gemm_dgrad_delta(d_den, p.wo, sv.ctx, dctx, ws, B, heads, Sq)
    else:
        ops.gemm(d_den, p.wo, dctx, b_layout=1)
    dx = torch.empty(Mq, H, dtype=F16, device=dev)
    if not cross:
        dqkv = torch.empty(Mq, 3 * H, dtype=F16, device=dev)
        ops.attn_bwd(sv.q, sv.q, dctx, sv.ctx, sv.lse2, dqkv, dqkv, ws, B, heads, Sq, Sq, q_col0=0, k_col0=H, v_col0=2 * H, dq_col0=0,
                     dk_col0=H, dv_col0=2 * H, key_bias=key_bias, kv_len=kv_len, drop=sv.drop_attn, delta_ready=fused_delta, pack=sv.pack,
                     dq_half=Experimental.dq16 and sv.pack is None)
        ops.colsum(dqkv, g.bqkv, inv_scale)
        ops.gemm(dqkv, sv.x16, g.wqkv, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv_scale,
                 k_splits=ops.wgrad_splits(3 * H, H, Mq))
        ops.gemm(dqkv, p.wqkv, dx, b_layout=1, epilogue=ops.EPI_ADD, aux=d_pre)
        return dx, None
    Mk, Hkv = B * Sk, p.wkv.shape[1]
    dq = torch.empty(Mq, H, dtype=F16, device=dev)
    dkv = torch.empty(Mk, 2 * H, dtype=F16, device=dev)
    ops.attn_bwd(sv.q, sv.kv, dctx, sv.ctx, sv.lse2, dq, dkv, ws, B, heads, Sq, Sk, q_col0=0, k_col0=0, v_col0=H, dq_col0=0, dk_col0=0,
                 dv_col0=H, key_bias=key_bias, kv_len=kv_len, drop=sv.drop_attn)
    ops.colsum(dq, g.bq, inv_scale)
    ops.colsum(dkv, g.bkv, inv_scale)
    ops.gemm(dq, sv.x16, g.wq, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv_scale, k_splits=ops.wgrad_splits(H, H, Mq))
    ops.gemm(dkv, sv.kv16, g.wkv, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, alpha=inv_scale, k_splits=ops.wgrad_splits(2 * H, Hkv, Mk))
    ops.gemm(dq, p.wq, dx, b_layout=1, epilogue=ops.EPI_ADD, aux=d_pre)
    dkv_src = torch.empty(Mk, Hkv, dtype=F16, device=dev)
    ops.gemm(dkv, p.wkv, dkv_src, b_layout=1)
    return dx, dkv_src
