"""Tensor-level wrappers over the C ABI: each function takes torch CUDA tensors, checks
dtype/contiguity, and enqueues one library call on torch's current stream.  PyTorch is used
for device memory and streams only; no arithmetic happens here."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as L
from .lib import (DT_F16, DT_F32, EPI_ADD, EPI_ATOMIC, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_BIAS_RES32,  # noqa: F401
                  EPI_DGELU, EPI_STORE)

Tensor = torch.Tensor


class Dropout:
    """(device seed tensor, site id, probability) of one dropout site; `None` anywhere means "no dropout"."""
    __slots__ = ("seed", "site", "p")

    def __init__(self, seed: Tensor, site: int, p: float):
        self.seed, self.site, self.p = seed, int(site), float(p)


def _drop_args(d):
    if d is None or d.p <= 0.0:
        return None, 0, 0.0
    return C.c_void_p(d.seed.data_ptr()), d.site, d.p


def _ptr(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: Tensor, dtype, name: str):
    if not t.is_cuda:
        raise L.B200Error(f"{name}: expected a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise L.B200Error(f"{name}: expected {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise L.B200Error(f"{name}: last dim must be contiguous")
    return t


def _dt(t: Tensor) -> int:
    return DT_F32 if t.dtype == torch.float32 else DT_F16


def cast_f32_to_bf16(src: Tensor, dst: Optional[Tensor] = None) -> Tensor:
    _req(src, torch.float32, "src")
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    L.check(L.load().b200_cast_f32_to_bf16(_ptr(src), _ptr(dst), src.numel(), _stream()), "b200_cast_f32_to_bf16")
    return dst


def gemm_bf16(a: Tensor, b: Tensor, out: Tensor, *, a_layout: int = 0, b_layout: int = 0, epilogue: int = EPI_STORE,
              bias: Optional[Tensor] = None, alpha: Optional[Tensor] = None, k_splits: int = 1) -> Tensor:
    """bf16-operand form of `gemm` (plain epilogues: STORE / BIAS forward, STORE dgrad, ATOMIC wgrad); out bf16 or fp32."""
    _req(a, torch.bfloat16, "A"), _req(b, torch.bfloat16, "B")
    if out.dtype not in (torch.bfloat16, torch.float32):
        raise L.B200Error(f"gemm_bf16: out must be bf16 or fp32, got {out.dtype}")
    M = a.shape[1] if a_layout else a.shape[0]
    K = a.shape[0] if a_layout else a.shape[1]
    N = b.shape[1] if b_layout else b.shape[0]
    rc = L.load().b200_gemm_bf16(_ptr(a), a.stride(0), a_layout, _ptr(b), b.stride(0), b_layout, M, N, K, epilogue, _ptr(bias), _ptr(out),
                                 out.stride(0), L.DT_F32 if out.dtype == torch.float32 else L.DT_BF16, _ptr(alpha), k_splits, _stream())
    L.check(rc, "b200_gemm_bf16")
    return out


class RowIndex:
    """Compaction of a [B, S] key tensor: the positions with key != ignore in row-major order (b200_heads_compact).  On an
    attention mask (ignore = 0) this IS the packed layout of SURVEY.md §8f rank 2: `start` = cu_seqlens [B+1], `idx` = flat
    position of every packed row, `ex` / `rank` = its sequence and position inside it, n = rows, max_n = longest sequence."""
    __slots__ = ("idx", "ex", "rank", "cnt", "start", "n", "max_n", "B", "S")

    def __init__(self, idx, ex, rank, cnt, start, n, max_n, B, S):
        self.idx, self.ex, self.rank, self.cnt, self.start, self.n, self.max_n, self.B, self.S = idx, ex, rank, cnt, start, n, max_n, B, S


def compact_rows(key: Tensor, ignore: int) -> RowIndex:
    """One host read (8 bytes: row count and longest list) sizes everything downstream."""
    if not key.is_cuda:
        raise L.B200Error("compact_rows: expected a CUDA tensor (the B200 path has no CPU fallback)")
    key = key.contiguous()
    if key.dtype != torch.int64:
        key = key.to(torch.int64)
    B, S = key.shape
    dev = key.device
    i32 = torch.int32
    tmp = torch.empty(B * S, dtype=i32, device=dev)
    cnt, start = torch.empty(B, dtype=i32, device=dev), torch.empty(B + 1, dtype=i32, device=dev)
    totals = torch.empty(2, dtype=i32, device=dev)
    idx, ex, rank = (torch.empty(B * S, dtype=i32, device=dev) for _ in range(3))
    L.check(L.load().b200_heads_compact(_ptr(key), int(ignore), B, S, _ptr(tmp), _ptr(cnt), _ptr(start), _ptr(totals), _ptr(idx), _ptr(ex),
                                        _ptr(rank), _stream()), "b200_heads_compact")
    n, max_n = totals.tolist()
    return RowIndex(idx[:n], ex[:n], rank[:n], cnt, start, int(n), int(max_n), B, S)


def gather_i64(key: Tensor, rows: RowIndex) -> Tensor:
    key = _req(key.contiguous().view(-1), torch.int64, "key")
    out = torch.empty(rows.n, dtype=torch.int64, device=key.device)
    L.check(L.load().b200_gather_i64(_ptr(key), _ptr(rows.idx), rows.n, _ptr(out), _stream()), "b200_gather_i64")
    return out


def unpack_rows(src: Tensor, rows: RowIndex, dst: Tensor) -> Tensor:
    """dst[idx[i]] = src[i] (fp32 rows): packed activations back into the padded [B*S, H] layout."""
    _req(src, torch.float32, "src"), _req(dst, torch.float32, "dst")
    L.check(L.load().b200_unpack_rows(_ptr(src), _ptr(rows.idx), rows.n, src.shape[1], _ptr(dst), _stream()), "b200_unpack_rows")
    return dst


def gemm(a: Tensor, b: Tensor, out: Tensor, *, a_layout: int = 0, b_layout: int = 0, epilogue: int = EPI_STORE,
         bias: Optional[Tensor] = None, aux: Optional[Tensor] = None, out2: Optional[Tensor] = None,
         alpha: Optional[Tensor] = None, k_splits: int = 1, drop: Optional["Dropout"] = None) -> Tensor:
    """out[M,N] = epilogue(A @ B^T).  a_layout/b_layout = 1 means the operand is stored transposed
    ([K,M] / [K,N] row-major) and is fed MN-major to the tensor core."""
    _req(a, torch.float16, "A"), _req(b, torch.float16, "B")
    M = a.shape[0] if a_layout == 0 else a.shape[1]
    K = a.shape[1] if a_layout == 0 else a.shape[0]
    N = b.shape[0] if b_layout == 0 else b.shape[1]
    Kb = b.shape[1] if b_layout == 0 else b.shape[0]
    if K != Kb or tuple(out.shape) != (M, N):
        raise L.B200Error(f"gemm: shape mismatch A{tuple(a.shape)}/{a_layout} B{tuple(b.shape)}/{b_layout} out{tuple(out.shape)}")
    if drop is not None and drop.p > 0.0:
        if a_layout or b_layout or out2 is not None or alpha is not None or k_splits != 1:
            raise L.B200Error("gemm: dropout is fused into the forward residual epilogues only")
        seed, site, p = _drop_args(drop)
        rc = L.load().b200_gemm_f16_drop(_ptr(a), a.stride(0), _ptr(b), b.stride(0), M, N, K, epilogue, _ptr(bias), _ptr(aux),
                                         aux.stride(0) if aux is not None else 0, _ptr(out), out.stride(0), _dt(out), seed, site, p, _stream())
        L.check(rc, "b200_gemm_f16_drop")
        return out
    rc = L.load().b200_gemm_f16(_ptr(a), a.stride(0), a_layout, _ptr(b), b.stride(0), b_layout, M, N, K, epilogue,
                                _ptr(bias), _ptr(aux), aux.stride(0) if aux is not None else 0, _ptr(out), out.stride(0),
                                _dt(out), _ptr(out2), out2.stride(0) if out2 is not None else 0, _ptr(alpha), k_splits,
                                _stream())
    L.check(rc, "b200_gemm_f16")
    return out


def gemm_resadd(a: Tensor, b: Tensor, out32: Tensor, bias: Tensor, *, drop: Optional["Dropout"] = None) -> Tensor:
    """out32[M,N] += dropout(A @ B^T + bias): `out32` holds the fp32 residual on entry and the pre-LayerNorm sum on return
    (b200_gemm_f16_resadd)."""
    _req(a, torch.float16, "A"), _req(b, torch.float16, "B"), _req(out32, torch.float32, "out32"), _req(bias, torch.float32, "bias")
    M, K, N = a.shape[0], a.shape[1], b.shape[0]
    if b.shape[1] != K or tuple(out32.shape) != (M, N) or bias.numel() != N:
        raise L.B200Error(f"gemm_resadd: shape mismatch A{tuple(a.shape)} B{tuple(b.shape)} out{tuple(out32.shape)} bias{tuple(bias.shape)}")
    seed, site, p = _drop_args(drop)
    rc = L.load().b200_gemm_f16_resadd(_ptr(a), a.stride(0), _ptr(b), b.stride(0), M, N, K, _ptr(bias), _ptr(out32), out32.stride(0),
                                       seed, site, p, _stream())
    L.check(rc, "b200_gemm_f16_resadd")
    return out32


def attn_bwd_delta_view(workspace: Tensor, B: int, heads: int, Sq: int) -> Tensor:
    """The [B, heads, Sq] fp32 row statistic at the head of the attention-backward workspace (b200_attn_bwd_delta_ptr)."""
    return workspace[:B * heads * Sq].view(B, heads, Sq)


def gemm_dgelu_colsum(d_dense: Tensor, w: Tensor, dact: Tensor, dz: Tensor, colsum: Tensor, col_alpha: Optional[Tensor] = None) -> Tensor:
    """dz = (d_dense @ W) * dact and colsum += col_alpha * dz.sum(0) in one kernel (W row-major [out, in] = [K, N])."""
    for t, n in ((d_dense, "d_dense"), (w, "W"), (dact, "dact"), (dz, "dz")):
        _req(t, torch.float16, n)
    _req(colsum, torch.float32, "colsum")
    M, K, N = d_dense.shape[0], d_dense.shape[1], w.shape[1]
    if w.shape[0] != K or tuple(dz.shape) != (M, N) or tuple(dact.shape) != (M, N) or colsum.numel() != N:
        raise L.B200Error(f"gemm_dgelu_colsum: shape mismatch d_dense{tuple(d_dense.shape)} W{tuple(w.shape)} dact{tuple(dact.shape)} colsum{tuple(colsum.shape)}")
    rc = L.load().b200_gemm_f16_dgelu_colsum(_ptr(d_dense), d_dense.stride(0), _ptr(w), w.stride(0), M, N, K, _ptr(dact), dact.stride(0),
                                             _ptr(dz), dz.stride(0), _ptr(colsum), _ptr(col_alpha), _stream())
    L.check(rc, "b200_gemm_f16_dgelu_colsum")
    return dz


def gemm_dgrad_delta(dy: Tensor, w: Tensor, ctx: Tensor, dctx: Tensor, workspace: Tensor, B: int, heads: int, Sq: int) -> Tensor:
    """dctx = dy @ W (the output projection's dgrad, W row-major [out, in]) and, in the same epilogue,
    delta[b,h,q] = <dctx row, ctx row> per head, written where `attn_bwd(..., delta_ready=True)` expects it."""
    for t, n in ((dy, "dy"), (w, "W"), (ctx, "ctx"), (dctx, "dctx")):
        _req(t, torch.float16, n)
    M, K, N = dy.shape[0], dy.shape[1], w.shape[1]
    if w.shape[0] != K or tuple(dctx.shape) != (M, N) or tuple(ctx.shape) != (M, N) or M != B * Sq or N != heads * 64:
        raise L.B200Error(f"gemm_dgrad_delta: shape mismatch dy{tuple(dy.shape)} W{tuple(w.shape)} ctx{tuple(ctx.shape)} dctx{tuple(dctx.shape)}")
    delta = attn_bwd_delta_view(_req(workspace, torch.float32, "workspace"), B, heads, Sq)
    rc = L.load().b200_gemm_f16_dgrad_delta(_ptr(dy), dy.stride(0), _ptr(w), w.stride(0), M, N, K, _ptr(ctx), ctx.stride(0), _ptr(dctx),
                                            dctx.stride(0), _ptr(delta), heads, Sq, _stream())
    L.check(rc, "b200_gemm_f16_dgrad_delta")
    return dctx


def wgrad_splits(M_out: int, N_in: int, K_tokens: int, sms: int = 148) -> int:
    """Split-K factor for a weight-gradient GEMM (2-CTA kernel: 256 x 256 pair tiles, sms/2 pairs): the smallest split
    count whose last round of the persistent grid is at least 90 % full, keeping >= 16 K blocks per split."""
    pairs = max(1, sms // 2)
    tiles = ((M_out + 255) // 256) * ((N_in + 255) // 256)
    kb = (K_tokens + 63) // 64
    best, best_eff = 1, 0.0
    for s in range(1, max(1, min(kb // 16, 64)) + 1):
        if (kb + s - 1) // s * (s - 1) >= kb:        # would leave an empty split
            continue
        units = tiles * s
        eff = units / (((units + pairs - 1) // pairs) * pairs)
        if eff > best_eff + 0.02:
            best, best_eff = s, eff
        if best_eff >= 0.9:
            break
    return best


def attn_fwd(q: Tensor, kv: Tensor, ctx: Tensor, B: int, heads: int, Sq: int, Sk: int, *, q_col0: int, k_col0: int,
             v_col0: int, key_bias: Optional[Tensor] = None, kv_len: Optional[Tensor] = None,
             lse2: Optional[Tensor] = None, drop: Optional["Dropout"] = None, pack: Optional[RowIndex] = None) -> Tensor:
    """`pack`: q / kv / ctx hold PACKED rows (sequence b = rows [cu[b], cu[b+1])); Sq = Sk = the padded length."""
    _req(q, torch.float16, "q"), _req(kv, torch.float16, "kv"), _req(ctx, torch.float16, "ctx")
    seed, site, p = _drop_args(drop)
    if pack is not None:
        if kv is not q or key_bias is not None or kv_len is not None:
            raise L.B200Error("attn_fwd: packed rows are for self-attention without an extra key mask")
        rc = L.load().b200_attn_fwd_varlen(_ptr(q), q.stride(0), q_col0, k_col0, v_col0, _ptr(pack.start), pack.n, _ptr(ctx), ctx.stride(0),
                                           _ptr(lse2), B, heads, Sq, seed, site, p, _stream())
        L.check(rc, "b200_attn_fwd_varlen")
        return ctx
    rc = L.load().b200_attn_fwd_drop(_ptr(q), q.stride(0), q_col0, _ptr(kv), kv.stride(0), k_col0, v_col0, _ptr(key_bias),
                                     _ptr(kv_len), _ptr(ctx), ctx.stride(0), _ptr(lse2), B, heads, Sq, Sk, seed, site, p, _stream())
    L.check(rc, "b200_attn_fwd")
    return ctx


def attn_probs(q: Tensor, k: Tensor, lse2: Tensor, B: int, heads: int, Sq: int, Sk: int, *, q_col0: int, k_col0: int,
               key_bias: Optional[Tensor] = None) -> Tensor:
    probs = torch.empty(B, heads, Sq, Sk, dtype=torch.float32, device=q.device)
    rc = L.load().b200_attn_probs(_ptr(q), q.stride(0), q_col0, _ptr(k), k.stride(0), k_col0, _ptr(key_bias), _ptr(lse2),
                                  _ptr(probs), B, heads, Sq, Sk, _stream())
    L.check(rc, "b200_attn_probs")
    return probs


def mask_to_bias(mask: Tensor):
    """[B,S] 0/1 mask -> (key_bias fp32 [B,S], kv_len int32 [B])."""
    B, S = mask.shape
    mask = mask.contiguous()
    code = {torch.int64: 0, torch.float32: 1, torch.int32: 2}.get(mask.dtype)
    if code is None:
        mask = mask.to(torch.int64)
        code = 0
    bias = torch.empty(B, S, dtype=torch.float32, device=mask.device)
    kv_len = torch.empty(B, dtype=torch.int32, device=mask.device)
    L.check(L.load().b200_mask_to_bias(_ptr(mask), code, _ptr(bias), _ptr(kv_len), B, S, _stream()), "b200_mask_to_bias")
    return bias, kv_len


def layernorm_fwd(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, *, y: Optional[Tensor] = None,
                  y32: Optional[Tensor] = None, mean: Optional[Tensor] = None, rstd: Optional[Tensor] = None) -> Tensor:
    rows, H = x.shape
    if y is None:
        y = torch.empty(rows, H, dtype=torch.float16, device=x.device)
    rc = L.load().b200_layernorm_fwd(_ptr(x), _dt(x), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(y32), _ptr(mean), _ptr(rstd),
                                     rows, H, float(eps), _stream())
    L.check(rc, "b200_layernorm_fwd")
    return y


def layernorm_bwd(dy: Tensor, x: Tensor, mean: Tensor, rstd: Tensor, gamma: Tensor, dx: Tensor, dgamma: Tensor,
                  dbeta: Tensor, *, dy2: Optional[Tensor] = None, dbias: Optional[Tensor] = None,
                  alpha: Optional[Tensor] = None, dx_drop: Optional[Tensor] = None, drop: Optional["Dropout"] = None) -> Tensor:
    rows, H = x.shape
    if drop is not None and drop.p > 0.0:
        seed, site, p = _drop_args(drop)
        rc = L.load().b200_layernorm_bwd_drop(_ptr(dy), _ptr(dy2), _ptr(x), _dt(x), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(dx), _ptr(dx_drop),
                                              _ptr(dgamma), _ptr(dbeta), _ptr(dbias), _ptr(alpha), rows, H, seed, site, p, _stream())
        L.check(rc, "b200_layernorm_bwd_drop")
        return dx
    rc = L.load().b200_layernorm_bwd(_ptr(dy), _ptr(dy2), _ptr(x), _dt(x), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(dx),
                                     _ptr(dgamma), _ptr(dbeta), _ptr(dbias), _ptr(alpha), rows, H, _stream())
    L.check(rc, "b200_layernorm_bwd")
    return dx


def embed_ln_fwd(ids, tt, pos, inputs_embeds, word, pos_tab, type_tab, gamma, beta, eps, rows, S, H, *, y=None, y32=None, drop=None):
    if y is None:
        y = torch.empty(rows, H, dtype=torch.float16, device=word.device)
    seed, site, p = _drop_args(drop)
    rc = L.load().b200_embed_ln_fwd_drop(_ptr(ids), _ptr(tt), _ptr(pos), _ptr(inputs_embeds), _ptr(word), _ptr(pos_tab),
                                         _ptr(type_tab), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(y32), rows, S, H, float(eps),
                                         seed, site, p, _stream())
    L.check(rc, "b200_embed_ln_fwd")
    return y


def embed_ln_bwd(dy, dy2, ids, tt, pos, word, pos_tab, type_tab, gamma, dword, dpos, dtype_tab, dgamma, dbeta, alpha, eps,
                 rows, S, H, drop=None, *, pad_id: Optional[int] = None, inputs_embeds=None, d_inputs_embeds=None):
    """`pad_id`: nn.Embedding's padding_idx (config.pad_token_id) — that row of dword receives nothing; None = no padding index.
    `inputs_embeds` / `d_inputs_embeds`: the forward ran on caller-provided embeddings (ids is None)."""
    seed, site, p = _drop_args(drop)
    rc = L.load().b200_embed_ln_bwd_drop(_ptr(dy), _ptr(dy2), _ptr(ids), _ptr(tt), _ptr(pos), _ptr(word), _ptr(pos_tab),
                                         _ptr(type_tab), _ptr(gamma), _ptr(dword), _ptr(dpos), _ptr(dtype_tab), _ptr(dgamma),
                                         _ptr(dbeta), _ptr(alpha), rows, S, H, float(eps), -1 if pad_id is None else int(pad_id),
                                         _ptr(inputs_embeds), _ptr(d_inputs_embeds), seed, site, p, _stream())
    L.check(rc, "b200_embed_ln_bwd")


def cls_head_fwd(h: Tensor, W: Tensor, b: Tensor, *, want_argmax: bool = False, drop=None):
    rows, H = h.shape
    Cn = W.shape[0]
    logits = torch.empty(rows, Cn, dtype=torch.float32, device=h.device)
    am = torch.empty(rows, dtype=torch.int32, device=h.device) if want_argmax else None
    seed, site, p = _drop_args(drop)
    rc = L.load().b200_cls_head_fwd_drop(_ptr(h), _ptr(W), _ptr(b), _ptr(logits), _ptr(am), rows, H, Cn, seed, site, p, _stream())
    L.check(rc, "b200_cls_head_fwd")
    return (logits, am) if want_argmax else logits


def ce_stats(logits: Tensor, labels: Tensor, stats: Tensor, class_weight: Optional[Tensor] = None) -> Tensor:
    rows, Cn = logits.shape
    L.check(L.load().b200_ce_stats(_ptr(logits), _ptr(labels), _ptr(class_weight), _ptr(stats), rows, Cn, _stream()),
            "b200_ce_stats")
    return stats


def cls_head_bwd(h, logits, labels, stats, W, dh, dW, db, *, class_weight=None, scale=None, drop=None):
    rows, H = h.shape
    seed, site, p = _drop_args(drop)
    rc = L.load().b200_cls_head_bwd_drop(_ptr(h), _ptr(logits), _ptr(labels), _ptr(class_weight), _ptr(stats), _ptr(W),
                                         _ptr(scale), _ptr(dh), _ptr(dW), _ptr(db), rows, H, W.shape[0], seed, site, p, _stream())
    L.check(rc, "b200_cls_head_bwd")
    return dh


def colsum(dy: Tensor, db: Tensor, alpha: Optional[Tensor] = None) -> Tensor:
    rows, cols = dy.shape
    L.check(L.load().b200_colsum(_ptr(dy), dy.stride(0), _ptr(db), _ptr(alpha), rows, cols, _stream()), "b200_colsum")
    return db


def cast_f32_to_f16(src: Tensor, dst: Tensor) -> Tensor:
    L.check(L.load().b200_cast_f32_to_f16(_ptr(src), _ptr(dst), src.numel(), _stream()), "b200_cast_f32_to_f16")
    return dst


def cast_f16_to_f32(src: Tensor, dst: Tensor) -> Tensor:
    L.check(L.load().b200_cast_f16_to_f32(_ptr(src), _ptr(dst), src.numel(), _stream()), "b200_cast_f16_to_f32")
    return dst


def scale_cast_grad(src: Tensor, dst: Tensor, scale: Tensor, amax_slot: Tensor, target: float = 1024.0) -> Tensor:
    rc = L.load().b200_scale_cast_grad(_ptr(src), _ptr(dst), src.numel(), float(target), _ptr(scale), _ptr(amax_slot), _stream())
    L.check(rc, "b200_scale_cast_grad")
    return dst


def attn_bwd_workspace(B: int, heads: int, Sq: int, device, rows: Optional[int] = None) -> Tensor:
    """`rows`: packed row count (the dQ accumulator then has that many rows; the row statistic keeps its padded layout)."""
    n = int(L.load().b200_attn_bwd_workspace(B, heads, Sq) if rows is None else L.load().b200_attn_bwd_workspace_varlen(B, heads, Sq, rows))
    return torch.empty(n // 4, dtype=torch.float32, device=device)


def attn_bwd(q: Tensor, kv: Tensor, dctx: Tensor, ctx: Tensor, lse2: Tensor, dq: Tensor, dkv: Tensor, workspace: Tensor,
             B: int, heads: int, Sq: int, Sk: int, *, q_col0: int, k_col0: int, v_col0: int, dq_col0: int, dk_col0: int,
             dv_col0: int, key_bias: Optional[Tensor] = None, kv_len: Optional[Tensor] = None, drop=None,
             delta_ready: bool = False, pack: Optional[RowIndex] = None, dq_half: bool = False) -> None:
    """`delta_ready`: the workspace already holds rowsum(dO o O) from `gemm_dgrad_delta`; skip that pass.  `pack`: packed rows.
    `dq_half`: dQ is accumulated as fp16 TMA reduce-adds straight into `dq` (no fp32 accumulator, memset or cast pass)."""
    for t, n in ((q, "q"), (kv, "kv"), (dctx, "dctx"), (ctx, "ctx"), (dq, "dq"), (dkv, "dkv")):
        _req(t, torch.float16, n)
    seed, site, p = _drop_args(drop)
    if pack is not None:
        if kv is not q or dkv is not dq or delta_ready or key_bias is not None or kv_len is not None:
            raise L.B200Error("attn_bwd: packed rows are for self-attention (one packed QKV / dQKV buffer, own row-statistic pass)")
        rc = L.load().b200_attn_bwd_varlen(_ptr(q), q.stride(0), q_col0, k_col0, v_col0, _ptr(dctx), dctx.stride(0), _ptr(ctx), ctx.stride(0),
                                           _ptr(pack.start), _ptr(pack.ex), _ptr(pack.rank), pack.n, _ptr(lse2), _ptr(workspace), _ptr(dq),
                                           dq.stride(0), dq_col0, dk_col0, dv_col0, B, heads, Sq, seed, site, p, _stream())
        L.check(rc, "b200_attn_bwd_varlen")
        return
    if delta_ready or dq_half:
        rc = L.load().b200_attn_bwd_ext(_ptr(q), q.stride(0), q_col0, _ptr(kv), kv.stride(0), k_col0, v_col0, _ptr(dctx), dctx.stride(0),
                                        _ptr(ctx), ctx.stride(0), _ptr(key_bias), _ptr(kv_len), _ptr(lse2), _ptr(workspace), _ptr(dq),
                                        dq.stride(0), dq_col0, _ptr(dkv), dkv.stride(0), dk_col0, dv_col0, B, heads, Sq, Sk, seed, site, p,
                                        (1 if delta_ready else 0) | (2 if dq_half else 0), _stream())
        L.check(rc, "b200_attn_bwd_ext")
        return
    rc = L.load().b200_attn_bwd_drop(_ptr(q), q.stride(0), q_col0, _ptr(kv), kv.stride(0), k_col0, v_col0, _ptr(dctx), dctx.stride(0),
                                     _ptr(ctx), ctx.stride(0), _ptr(key_bias), _ptr(kv_len), _ptr(lse2), _ptr(workspace), _ptr(dq),
                                     dq.stride(0), dq_col0, _ptr(dkv), dkv.stride(0), dk_col0, dv_col0, B, heads, Sq, Sk, seed, site, p,
                                     _stream())
    L.check(rc, "b200_attn_bwd")


def grad_sumsq(g: Tensor, sumsq: Tensor) -> Tensor:
    L.check(L.load().b200_grad_sumsq(_ptr(g), g.numel(), _ptr(sumsq), _stream()), "b200_grad_sumsq")
    return sumsq


def clip_coef(sumsq: Tensor, coef: Tensor, max_norm: float, grad_mult: float = 1.0) -> Tensor:
    L.check(L.load().b200_clip_coef(_ptr(sumsq), float(max_norm), float(grad_mult), _ptr(coef), _stream()), "b200_clip_coef")
    return coef


def clip_coef_scaled(sumsq: Tensor, coef: Tensor, max_norm: float, grad_mult: float, loss_scale: Tensor, state: Tensor, *,
                     growth_interval: int = 2000, backoff: float = 0.5, growth: float = 2.0, min_scale: float = 1.0,
                     max_scale: float = 2.0 ** 24) -> Tensor:
    """clip_coef + dynamic loss scaling on the device: `loss_scale` = fp32 [2] {scale, 1/scale}, `state` = fp32 [2]
    {consecutive finite steps, skipped steps}."""
    _req(loss_scale, torch.float32, "loss_scale")
    _req(state, torch.float32, "state")
    L.check(L.load().b200_clip_coef_scaled(_ptr(sumsq), float(max_norm), float(grad_mult), _ptr(coef), _ptr(loss_scale), _ptr(state),
                                           int(growth_interval), float(backoff), float(growth), float(min_scale), float(max_scale),
                                           _stream()), "b200_clip_coef_scaled")
    return coef


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, p16: Optional[Tensor], *, lr: float, beta1: float = 0.9,
               beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.0, step: int = 1,
               coef: Optional[Tensor] = None) -> None:
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    rc = L.load().b200_adamw_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p16), p.numel(), lr, beta1, beta2, eps, weight_decay,
                                  bc1, bc2, _ptr(coef), _stream())
    L.check(rc, "b200_adamw_step")


def unscale_cast_grad(src: Tensor, dst: Tensor, scale: Optional[Tensor]) -> Tensor:
    """dst(fp32) = src(fp16) * scale[1]."""
    L.check(L.load().b200_unscale_cast_grad(_ptr(src), _ptr(dst), src.numel(), _ptr(scale), _stream()), "b200_unscale_cast_grad")
    return dst


def ponet_mix_fwd(proj: Tensor, seg_ids: Tensor, out: Tensor, B: int, S: int, heads: int, nseg: int, *,
                  key_bias: Optional[Tensor] = None) -> Tensor:
    """PoNet pooling mixer on the packed projections [B*S, 5H] = [Q | K | O | Sg | Lc] (oracle/ponet_oracle.py).
    Returns the workspace the kernels filled (global vector, segment maxima, ...), which the backward consumes."""
    _req(proj, torch.float16, "proj"), _req(out, torch.float16, "out")
    H = heads * 64
    nbytes = int(L.load().b200_ponet_workspace(B, S, H, heads, nseg))
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=proj.device)
    rc = L.load().b200_ponet_mix_fwd(_ptr(proj), proj.stride(0), _ptr(key_bias), _ptr(seg_ids), _ptr(ws), _ptr(out), B, S, H, heads, nseg,
                                     _stream())
    L.check(rc, "b200_ponet_mix_fwd")
    return ws


def ponet_mix_bwd(proj: Tensor, dout: Tensor, seg_ids: Tensor, fwd_ws: Tensor, dproj: Tensor, B: int, S: int, heads: int, nseg: int, *,
                  key_bias: Optional[Tensor] = None) -> Tensor:
    _req(proj, torch.float16, "proj"), _req(dout, torch.float16, "dout"), _req(dproj, torch.float16, "dproj")
    H = heads * 64
    nbytes = int(L.load().b200_ponet_bwd_workspace(B, S, H, heads, nseg))
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=proj.device)
    rc = L.load().b200_ponet_mix_bwd(_ptr(proj), proj.stride(0), _ptr(dout), _ptr(key_bias), _ptr(seg_ids), _ptr(fwd_ws), _ptr(ws), _ptr(dproj),
                                     dproj.stride(0), B, S, H, heads, nseg, _stream())
    L.check(rc, "b200_ponet_mix_bwd")
    return dproj


def set_hyper(hyper: Tensor, *, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.0,
              step: int = 1) -> None:
    rc = L.load().b200_set_hyper(_ptr(hyper), lr, beta1, beta2, eps, weight_decay, 1.0 - beta1 ** step, 1.0 - beta2 ** step, _stream())
    L.check(rc, "b200_set_hyper")


def adamw_step_dev(p: Tensor, g: Tensor, m: Tensor, v: Tensor, p16: Optional[Tensor], hyper: Tensor, coef: Optional[Tensor],
                   zero_grad: bool = False) -> None:
    """`zero_grad=True`: the kernel clears `g` after reading it (no separate fill pass before the next backward)."""
    fn = L.load().b200_adamw_step_dev_zero if zero_grad else L.load().b200_adamw_step_dev
    rc = fn(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p16), p.numel(), _ptr(hyper), _ptr(coef), _stream())
    L.check(rc, "b200_adamw_step_dev")
