"""Loss heads of the topic-segmentation wrapper on the labelled [BOS] rows, backed by libb200enc.so.

Drop-in surface — same names, arguments and returns as the reference (emnlp2023-topic_segmentation/src/models/modules/):
  * `LossCalculator(config)`: `.classifier` (Linear H -> num_labels), `.tssp_model.classifier` (Linear H -> num_tssp_labels);
    `forward(sequence_output, labels, extract_eop_segment_ids, eop_index_for_aggregate_batch_eop_features, sent_token_mask=None,
    sent_pair_orders=None, da_example_flag=False)` -> `(loss, logits, eop_pair_cos_sim)`        loss_calculator.py:8-73
  * the wrapper's two-view logic (bert_for_ts.py:35-113) is `topic_segmentation_forward` below.
What runs where: every sum over rows / features is a kernel of csrc/heads.cuh (compaction of the labelled positions, row
gather, Linear + CE / focal over all positions, adjacent-pair cosine, the CSSL similarity matrix or index lists, the TSSP
head, and all their backward passes); torch is left with 0-d scalar glue (adding the weighted loss terms) and allocation.
Host synchronisation: ONE 8-byte read per compaction (row count / longest example — they size the outputs the reference
returns, `[B, max_n]`); the compaction depends on the labels only.  `cl_anchor_level == "eop_list"` additionally reads the
topic ids back, because the reference draws its fall-back indices with Python's `random` on the host (cssl.py:152-166) and the
same draws, in the same order, are made here.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import random
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import lib as L
from .lib import B200Error

Tensor = torch.Tensor
F32, I32 = torch.float32, torch.int32


def _p(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(rc: int, what: str) -> None:
    L.check(rc, what)


from .ops import RowIndex, compact_rows as compact  # noqa: E402  (the compaction of the labelled positions doubles as the padding packer)


def gather_keys(key: Tensor, rows: RowIndex) -> Tensor:
    out = torch.empty(rows.n, dtype=I32, device=key.device)
    key = key.contiguous().to(torch.int64)
    _chk(L.load().b200_heads_gather_keys(_p(key), _p(rows.idx), rows.n, _p(out), _st()), "b200_heads_gather_keys")
    return out


def eop_list_indices(seg: Sequence[int], k_pos: int, k_neg: int, rng=random) -> Tuple[List[List[int]], List[List[int]]]:
    """Index choice of cssl.py:128-166 ("eop_list"): for row i of topic [start, end]: positives = the k_pos rows before i, falling
    back to a random row of [start, end) (or `end` itself when the topic has a single row) once they leave the topic; negatives =
    the k_neg rows after `end`, falling back to a random row of (end, last] (or of the first topic when nothing follows).  The
    draws are made row by row, positives first, with `rng.choice` — the reference's order, so a shared seed gives its lists."""
    n = len(seg)
    starts, ends = {}, {}
    for i, t in enumerate(seg):
        starts.setdefault(t, i)
        ends[t] = i
    first_topic = list(range(starts[seg[0]], starts[seg[0] + 1])) if (seg[0] + 1) in starts else []
    pos: List[List[int]] = [[] for _ in range(k_pos)]
    neg: List[List[int]] = [[] for _ in range(k_neg)]
    for i, t in enumerate(seg):
        s, e = starts[t], ends[t]
        pool = list(range(s, e)) or [e]
        j = i
        for k in range(k_pos):
            j -= 1
            if j < s:
                j = rng.choice(pool)
            pos[k].append(j)
        pool = list(range(e + 1, n)) or first_topic
        j = e
        for k in range(k_neg):
            j += 1
            if j >= n:
                j = rng.choice(pool)
            neg[k].append(j)
    return pos, neg


class _HeadsFn(torch.autograd.Function):
    """LossCalculator.forward as one autograd node: (loss, logits, cos) from the encoder output and the head weights."""

    @staticmethod
    def forward(ctx, h, cls_w, cls_b, tssp_w, tssp_b, labels, seg_ids, eop_index, stm, spo, cfg, da_example, rng):
        so = L.load()
        if not h.is_cuda:
            raise B200Error("heads: expected CUDA tensors (the B200 path has no CPU fallback)")
        B, S, H = h.shape
        dev = h.device
        h32 = h.detach().to(F32).contiguous()
        hf = h32.view(B * S, H)
        labels = labels.contiguous().to(torch.int64)
        rows = compact(labels, -100)
        n = rows.n
        st = _st()
        loss = torch.zeros(1, dtype=F32, device=dev)
        sv = dict(B=B, S=S, H=H, rows=rows, hf=hf, labels=labels, cfg=cfg, da=da_example, parts=[])
        # ---- rows, their normalised copies, the adjacent-pair cosine (utils.py:116-138) — always returned
        R = torch.empty(n, H, dtype=F32, device=dev)
        Rn = torch.empty(n, H, dtype=F32, device=dev)
        rinv = torch.empty(n, dtype=F32, device=dev)
        _chk(so.b200_heads_gather_rows(_p(hf), _p(rows.idx), n, H, _p(R), st), "b200_heads_gather_rows")
        _chk(so.b200_heads_normalize(_p(R), n, H, 1e-8, _p(Rn), _p(rinv), st), "b200_heads_normalize")
        lab_rows = gather_keys(labels, rows)
        cos = torch.full((B, rows.max_n), -100.0, dtype=F32, device=dev)
        cos_rows = torch.empty(n, dtype=F32, device=dev)
        temp = float(cfg.ts_score_predictor_cos_temp)
        if n:
            _chk(so.b200_heads_pair_cos_fwd(_p(Rn), _p(rows.ex), _p(rows.rank), _p(rows.start), _p(rows.cnt), n, H, temp, _p(cos_rows), _p(cos),
                                            max(rows.max_n, 1), st), "b200_heads_pair_cos_fwd")
        sv.update(Rn=Rn, rinv=rinv, lab_rows=lab_rows, cos_rows=cos_rows, temp=temp)
        # ---- the topic-segmentation loss proper (loss_calculator.py:41-50)
        if cfg.ts_score_predictor == "lt":
            Cn = cls_w.shape[0]
            w32, b32 = cls_w.detach().to(F32).contiguous(), cls_b.detach().to(F32).contiguous()
            logits = torch.empty(B, S, Cn, dtype=F32, device=dev)
            _chk(so.b200_heads_cls_fwd(_p(hf), _p(w32), _p(b32), _p(logits), None, B * S, H, Cn, st), "b200_heads_cls_fwd")
            cw = None
            if cfg.weight_label_zero != 0.5:                       # utils.py:173-182: class weights [w0, 1 - w0]
                cw = torch.tensor([cfg.weight_label_zero, 1.0 - cfg.weight_label_zero], dtype=F32, device=dev)
            stats = torch.zeros(3, dtype=F32, device=dev)
            gamma = float(cfg.focal_loss_gamma)
            _chk(so.b200_heads_focal_stats(_p(logits), _p(labels), _p(cw), gamma, B * S, Cn, _p(stats), st), "b200_heads_focal_stats")
            ce = stats[0:1] / stats[1:2]
            ts = ce if gamma == 0 else ce * (stats[2:3] / float(B * S))            # FocalLoss: mean_i (1 - p_i)^gamma x mean CE (utils.py:141-170)
            sv["parts"].append(("lt", dict(w32=w32, logits=logits, cw=cw, stats=stats, gamma=gamma, Cn=Cn)))
        elif cfg.ts_score_predictor == "cos":
            stats = torch.zeros(1, dtype=F32, device=dev)
            _chk(so.b200_heads_bce_fwd(_p(cos_rows), _p(lab_rows), n, _p(stats), st), "b200_heads_bce_fwd")
            total = B * rows.max_n
            # the reference feeds the PADDED matrix to BCEWithLogits: a padding cell has logit -100 and target -100.0, i.e. the
            # constant term max(x,0) - x*y + log1p(exp(-|x|)) = -10000 (kept: loss_calculator.py:46-49)
            pad_term = -(-100.0) * (-100.0) + float(torch.log1p(torch.exp(torch.tensor(-100.0))))
            ts = (stats + (total - n) * pad_term) / max(total, 1)
            logits = torch.sigmoid(cos)
            sv["parts"].append(("cos", dict(total=total)))
        else:
            raise ValueError(cfg.ts_score_predictor)
        loss = loss + float(cfg.ts_loss_weight) * ts
        # ---- CSSL on the anchor view (loss_calculator.py:54-64, cssl.py:224-273)
        if not da_example and cfg.cl_loss_weight != 0 and n > 2:
            seg = torch.empty(n, dtype=I32, device=dev)
            _chk(so.b200_heads_topic_ids(_p(lab_rows), _p(rows.ex), n, _p(seg), st), "b200_heads_topic_ids")
            seg_ids64 = seg_ids.contiguous().to(torch.int64)
            slots = compact(eop_index, 0)
            if slots.n != n:
                raise B200Error(f"CSSL: {slots.n} gathered [BOS] features for {n} labelled rows (the reference pairs them one to one)")
            slot_id = gather_keys(eop_index, slots)
            Fm = torch.empty(n, H, dtype=F32, device=dev)
            Fn = torch.empty(n, H, dtype=F32, device=dev)
            finv = torch.empty(n, dtype=F32, device=dev)
            _chk(so.b200_heads_segmax_fwd(_p(hf), _p(seg_ids64), _p(slots.ex), _p(slot_id), n, S, H, _p(Fm), st), "b200_heads_segmax_fwd")
            _chk(so.b200_heads_normalize(_p(Fm), n, H, 1e-8, _p(Fn), _p(finv), st), "b200_heads_normalize")
            cl = torch.zeros(1, dtype=F32, device=dev)
            part = dict(seg=seg, seg_ids=seg_ids64, slots=slots, slot_id=slot_id, Fm=Fm, Fn=Fn, finv=finv, temp=float(cfg.cl_temp))
            if cfg.cl_anchor_level == "eop_matrix":
                E = torch.empty(n, n, dtype=F32, device=dev)
                num, den, coef = (torch.empty(n, dtype=F32, device=dev) for _ in range(3))
                _chk(so.b200_heads_cssl_matrix_fwd(_p(Fn), _p(seg), n, H, part["temp"], float(cfg.cl_loss_weight), _p(E), _p(num), _p(den), _p(coef),
                                                   _p(cl), st), "b200_heads_cssl_matrix_fwd")
                part.update(kind="matrix", E=E, num=num, den=den, coef=coef)
                sv["parts"].append(("cssl", part))
                loss = loss + cl
            elif cfg.cl_anchor_level == "eop_list":
                seg_host = seg.tolist()                                   # host draw of the fall-back indices, as the reference
                if seg_host[-1] != 0:                                     # at least two topics (cssl.py:264-266)
                    pos, neg = eop_list_indices(seg_host, int(cfg.cl_positive_k), int(cfg.cl_negative_k), rng)
                    pos_t = torch.tensor(pos, dtype=I32, device=dev).contiguous()
                    neg_t = torch.tensor(neg, dtype=I32, device=dev).contiguous()
                    kp, kn = pos_t.shape[0], neg_t.shape[0]
                    gw = torch.empty(kp + kn, n, dtype=F32, device=dev)
                    _chk(so.b200_heads_cssl_list_fwd(_p(Fn), _p(pos_t), _p(neg_t), kp, kn, n, H, part["temp"], float(cfg.cl_loss_weight), _p(cl),
                                                     _p(gw), st), "b200_heads_cssl_list_fwd")
                    part.update(kind="list", pos=pos_t, neg=neg_t, gw=gw, kp=kp, kn=kn)
                    sv["parts"].append(("cssl", part))
                    loss = loss + cl
            else:
                raise ValueError(f"cl_anchor_level {cfg.cl_anchor_level!r} is not implemented (eop_matrix and eop_list are)")
        # ---- TSSP on the augmented view (loss_calculator.py:66-71, tssp.py:17-35)
        if da_example and cfg.tssp_loss_weight != 0:
            ra, rb = compact(stm, -100), compact(spo, -100)
            if ra.n != rb.n:
                raise B200Error(f"TSSP: {ra.n} sentence rows for {rb.n} order labels")
            tgt = gather_keys(spo, rb)
            R2 = torch.empty(ra.n, H, dtype=F32, device=dev)
            _chk(so.b200_heads_gather_rows(_p(hf), _p(ra.idx), ra.n, H, _p(R2), st), "b200_heads_gather_rows")
            tw32, tb32 = tssp_w.detach().to(F32).contiguous(), tssp_b.detach().to(F32).contiguous()
            C3 = tw32.shape[0]
            probs = torch.empty(ra.n, C3, dtype=F32, device=dev)
            tstats = torch.zeros(1, dtype=F32, device=dev)
            _chk(so.b200_heads_rows_ce_fwd(_p(R2), _p(tw32), _p(tb32), _p(tgt), ra.n, H, C3, _p(probs), _p(tstats), st), "b200_heads_rows_ce_fwd")
            # tssp.py:34 returns tssp_loss_weight * CE and loss_calculator.py:71 multiplies by tssp_loss_weight again (kept)
            wt = float(cfg.tssp_loss_weight) ** 2
            loss = loss + wt * tstats / max(ra.n, 1)
            sv["parts"].append(("tssp", dict(ra=ra, tgt=tgt, R2=R2, tw32=tw32, probs=probs, C3=C3, scale=wt / max(ra.n, 1))))
        ctx.sv = sv
        ctx.mark_non_differentiable(logits, cos)
        return loss.reshape(()), logits, cos

    @staticmethod
    def backward(ctx, g_loss, g_logits, g_cos):
        so = L.load()
        sv = ctx.sv
        B, S, H, rows, hf, cfg = sv["B"], sv["S"], sv["H"], sv["rows"], sv["hf"], sv["cfg"]
        n, dev, st = rows.n, hf.device, _st()
        g = g_loss.detach().to(F32).reshape(1).contiguous()           # upstream gradient of the scalar loss, read on the device
        dh = torch.zeros(B * S, H, dtype=F32, device=dev)
        dRn = torch.zeros(n, H, dtype=F32, device=dev)                # gradient wrt the normalised labelled rows (cosine paths)
        d_cls_w = d_cls_b = d_tssp_w = d_tssp_b = None
        for kind, p in sv["parts"]:
            if kind == "lt":
                d_cls_w, d_cls_b = torch.zeros_like(p["w32"]), torch.zeros(p["Cn"], dtype=F32, device=dev)
                _chk(so.b200_heads_cls_bwd(_p(hf), _p(p["logits"]), _p(sv["labels"]), _p(p["cw"]), _p(p["stats"]), _p(p["w32"]), p["gamma"],
                                           float(cfg.ts_loss_weight), _p(g), B * S, H, p["Cn"], _p(dh), _p(d_cls_w), _p(d_cls_b), st), "b200_heads_cls_bwd")
            elif kind == "cos":
                gc = torch.zeros(n, dtype=F32, device=dev)
                _chk(so.b200_heads_bce_bwd(_p(sv["cos_rows"]), _p(sv["lab_rows"]), n, float(cfg.ts_loss_weight) / max(p["total"], 1), _p(g), _p(gc), st),
                     "b200_heads_bce_bwd")
                _chk(so.b200_heads_pair_cos_bwd(_p(sv["Rn"]), _p(rows.ex), _p(rows.rank), _p(rows.start), _p(rows.cnt), _p(gc), n, H, sv["temp"],
                                                _p(dRn), st), "b200_heads_pair_cos_bwd")
            elif kind == "cssl":
                dFn = torch.zeros(n, H, dtype=F32, device=dev)
                if p["kind"] == "matrix":
                    _chk(so.b200_heads_cssl_matrix_bwd(_p(p["Fn"]), _p(p["seg"]), _p(p["E"]), _p(p["num"]), _p(p["den"]), _p(p["coef"]), n, H, p["temp"],
                                                       _p(dFn), st), "b200_heads_cssl_matrix_bwd")
                else:
                    _chk(so.b200_heads_cssl_list_bwd(_p(p["Fn"]), _p(p["pos"]), _p(p["neg"]), _p(p["gw"]), p["kp"], p["kn"], n, H, _p(dFn), st),
                         "b200_heads_cssl_list_bwd")
                dF = torch.zeros(n, H, dtype=F32, device=dev)
                _chk(so.b200_heads_normalize_bwd(_p(p["Fn"]), _p(p["finv"]), _p(dFn), n, H, 1.0, _p(g), _p(dF), st), "b200_heads_normalize_bwd")
                _chk(so.b200_heads_segmax_bwd(_p(hf), _p(p["seg_ids"]), _p(p["slots"].ex), _p(p["slot_id"]), _p(p["Fm"]), _p(dF), n, S, H, 1.0, _p(dh), st),
                     "b200_heads_segmax_bwd")
            elif kind == "tssp":
                ra = p["ra"]
                dR2 = torch.zeros(ra.n, H, dtype=F32, device=dev)
                d_tssp_w, d_tssp_b = torch.zeros_like(p["tw32"]), torch.zeros(p["C3"], dtype=F32, device=dev)
                _chk(so.b200_heads_rows_ce_bwd(_p(p["R2"]), _p(p["tw32"]), _p(p["tgt"]), _p(p["probs"]), ra.n, H, p["C3"], p["scale"], _p(g), _p(dR2),
                                               _p(d_tssp_w), _p(d_tssp_b), st), "b200_heads_rows_ce_bwd")
                _chk(so.b200_heads_scatter_rows(_p(dR2), _p(ra.idx), ra.n, H, 1.0, _p(dh), st), "b200_heads_scatter_rows")
        if any(k == "cos" for k, _ in sv["parts"]) and n:
            dR = torch.zeros(n, H, dtype=F32, device=dev)
            _chk(so.b200_heads_normalize_bwd(_p(sv["Rn"]), _p(sv["rinv"]), _p(dRn), n, H, 1.0, None, _p(dR), st), "b200_heads_normalize_bwd")
            _chk(so.b200_heads_scatter_rows(_p(dR), _p(rows.idx), n, H, 1.0, _p(dh), st), "b200_heads_scatter_rows")
        ctx.sv = None
        return (dh.view(B, S, H), d_cls_w, d_cls_b, d_tssp_w, d_tssp_b) + (None,) * 8


class TSSP(nn.Module):
    """tssp.py:8-35 — parameter holder (Linear H -> num_tssp_labels); the arithmetic runs inside LossCalculator's kernels."""

    def __init__(self, config):
        super().__init__()
        self.classifier = nn.Linear(config.hidden_size, config.num_tssp_labels)


class LossCalculator(nn.Module):
    """Drop-in for models.modules.loss_calculator.LossCalculator (same constructor argument, forward signature and return)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.num_labels = config.num_labels
        self.classifier = nn.Linear(config.hidden_size, config.num_labels)
        self.tssp_model = TSSP(config)
        self.rng = random                      # the reference draws eop_list fall-backs from the global `random` module

    def forward(self, sequence_output, labels, extract_eop_segment_ids, eop_index_for_aggregate_batch_eop_features, sent_token_mask=None,
                sent_pair_orders=None, da_example_flag=False):
        c = self.config
        return _HeadsFn.apply(sequence_output, self.classifier.weight, self.classifier.bias, self.tssp_model.classifier.weight,
                              self.tssp_model.classifier.bias, labels, extract_eop_segment_ids, eop_index_for_aggregate_batch_eop_features,
                              sent_token_mask, sent_pair_orders, c, bool(da_example_flag), self.rng)


def topic_segmentation_forward(bert, dropout, loss_calculator, config, input_ids, attention_mask, token_type_ids, labels,
                               extract_eop_segment_ids, eop_index_for_aggregate_batch_eop_features, sent_token_mask, sent_pair_orders):
    """bert_for_ts.py:35-113: anchor view -> heads; augmented view (when do_da_ts / do_tssp) -> heads with da_example_flag; summed
    loss, logits stacked [B, 2, S, C] (view 1 = a copy of view 0 without augmentation), the anchor view's pair cosines."""
    def enc(v):
        return dropout(bert(input_ids[:, v], attention_mask=attention_mask[:, v], head_mask=None, token_type_ids=token_type_ids[:, v],
                            position_ids=None, inputs_embeds=None, output_attentions=None, output_hidden_states=None, return_dict=False)[0])
    loss, logits0, cos = loss_calculator(enc(0), labels[:, 0], extract_eop_segment_ids[:, 0], eop_index_for_aggregate_batch_eop_features[:, 0])
    logits1 = logits0
    if config.do_da_ts or config.do_tssp:
        da_loss, logits1, _ = loss_calculator(enc(1), labels[:, 1], extract_eop_segment_ids[:, 1], eop_index_for_aggregate_batch_eop_features[:, 1],
                                              sent_token_mask=sent_token_mask[:, 1], sent_pair_orders=sent_pair_orders[:, 1], da_example_flag=True)
        loss = loss + da_loss
    return loss, torch.stack([logits0, logits1], 1), cos
