"""CPU oracle for the PoNet encoder layer (TEST INFRASTRUCTURE ONLY) — **PARITY UNPINNED**.

The reference's PoNet implementation is not in its tree: `alimeeting4mug/src/models/modeling_ponet.py:28` imports
`PoNetModel` from `modelscope==1.1.0` (`alimeeting4mug/requirements.txt:56`), which is not installed in this image and
cannot be fetched (no network).  This file therefore restates the *published* algorithm (PoNet, ICLR 2022, Tan et al.;
as summarised in SURVEY.md §8c "PoNet restatement") and is anchored only on the reference's call site
(`modeling_ponet.py:68-79`: `self.ponet(input_ids, attention_mask, token_type_ids, segment_ids, ...)`) and on the data it
feeds (`ponet_topic_segmentation.py:564-596,638,668`: `segment_ids` = 0 for CLS, 1..k per sentence, k+1 on padding —
monotone non-decreasing along the sequence; position table tiled to 4096, `:466-482`).  No vector of the reference pins
it: GPU parity for PoNet is "CUDA path == this restatement", nothing more, until modelscope becomes available.

Per layer, x [B,S,H], key-padding mask [B,S] (1 keep), segment_ids [B,S]:
    Q = dense_q(x), K = dense_k(x) (V == K), O = dense_o(x), Sg = dense_segment(x), Lc = dense_local(x)
    global : per head  qbar = masked-mean_s(Q);  a = softmax_s(qbar . K_s / sqrt(d) + mask);  g = sum_s a_s K_s    [B,h,d]
    segment: Sg := -1e4 on padding; seg_s = max over { t : segment_ids[t] == segment_ids[s] } Sg_t
    local  : Lc := -1e4 on padding; loc_s = max(Lc_{s-1}, Lc_s, Lc_{s+1})          (max_pool1d k=3, stride 1, pad 1)
    mix    : out_s = (g + seg_s) o O_s + loc_s, zeroed on padding
    then PoNetSelfOutput = LayerNorm(dense(out) + x) and the standard BERT feed-forward block.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from . import bert_oracle as O

Tensor = torch.Tensor
NEG = -10000.0


def ponet_mixer(q: Tensor, k: Tensor, o: Tensor, sg: Tensor, lc: Tensor, mask01: Optional[Tensor], segment_ids: Tensor,
                n_heads: int) -> Tensor:
    """All inputs [B,S,H]; returns the mixed context [B,S,H]."""
    B, S, H = q.shape
    d = H // n_heads
    keep = torch.ones(B, S, dtype=torch.bool) if mask01 is None else mask01.bool()
    kf = keep[:, :, None].to(q.dtype)
    # global: attention of the mean query over the keys (values == keys)
    qbar = (q * kf).sum(1) / kf.sum(1).clamp_min(1.0)                                    # [B,H]
    qh = qbar.view(B, n_heads, d)
    kh = k.view(B, S, n_heads, d)
    att = torch.einsum("bhd,bshd->bhs", qh, kh) / math.sqrt(d)
    att = att.masked_fill(~keep[:, None, :], float("-inf"))
    a = torch.softmax(att, dim=-1)
    g = torch.einsum("bhs,bshd->bhd", a, kh).reshape(B, 1, H)
    # segment max-pooling (scatter-max by segment id, gathered back)
    sgm = sg.masked_fill(~keep[:, :, None], NEG)
    nseg = int(segment_ids.max()) + 1
    idx = segment_ids[:, :, None].expand(B, S, H)
    pooled = torch.full((B, nseg, H), float("-inf"), dtype=sg.dtype).scatter_reduce(1, idx, sgm, reduce="amax", include_self=True)
    seg = pooled.gather(1, idx)
    # local max-pooling, window 3
    lcm = lc.masked_fill(~keep[:, :, None], NEG)
    loc = torch.nn.functional.max_pool1d(lcm.transpose(1, 2), kernel_size=3, stride=1, padding=1).transpose(1, 2)
    out = (g + seg) * o + loc
    return out * kf


def ponet_layer(sd: Dict[str, Tensor], p: str, cfg: O.OracleConfig, x: Tensor, mask01: Optional[Tensor], segment_ids: Tensor) -> Tensor:
    a = p + "attention.self."
    proj = {n: O.linear(x, sd[a + f"dense_{n}.weight"], sd[a + f"dense_{n}.bias"]) for n in ("q", "k", "o", "segment", "local")}
    ctx = ponet_mixer(proj["q"], proj["k"], proj["o"], proj["segment"], proj["local"], mask01, segment_ids, cfg.num_attention_heads)
    y = O.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
    y = O.layer_norm(y + x, sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"], cfg.layer_norm_eps)
    return O.ffn_block(sd, p, cfg, y)


def ponet_model(sd: Dict[str, Tensor], cfg: O.OracleConfig, input_ids: Tensor, attention_mask: Optional[Tensor],
                token_type_ids: Optional[Tensor], segment_ids: Tensor, position_ids: Optional[Tensor] = None) -> List[Tensor]:
    """Embeddings (BERT-style, position table as provided) + L PoNet layers; returns all hidden states."""
    x = O.embeddings(sd, cfg, input_ids, token_type_ids, position_ids)
    hs = [x]
    for i in range(cfg.num_hidden_layers):
        x = ponet_layer(sd, f"encoder.layer.{i}.", cfg, x, attention_mask, segment_ids)
        hs.append(x)
    return hs


def random_state_dict(cfg: O.OracleConfig, seed: int = 0) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    H, I = cfg.hidden_size, cfg.intermediate_size

    def n(*shape, s=0.02):
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
        for nm in ("q", "k", "o", "segment", "local"):
            sd[p + f"attention.self.dense_{nm}.weight"] = n(H, H, s=0.05)
            sd[p + f"attention.self.dense_{nm}.bias"] = n(H)
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
    return sd


def synth_segments(B: int, S: int, seed: int, pad_from: Optional[List[int]] = None):
    """segment_ids as the reference driver builds them: 0 for CLS, runs of ~U[8,40] tokens numbered 1..k, k+1 on padding."""
    g = torch.Generator().manual_seed(seed)
    seg = torch.zeros(B, S, dtype=torch.long)
    mask = torch.ones(B, S, dtype=torch.long)
    for b in range(B):
        end = S if pad_from is None else pad_from[b]
        s, k = 1, 0
        while s < end:
            k += 1
            ln = int(torch.randint(8, 41, (1,), generator=g))
            seg[b, s:min(end, s + ln)] = k
            s += ln
        seg[b, end:] = k + 1
        mask[b, end:] = 0
    return seg, mask
