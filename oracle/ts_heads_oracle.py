"""CPU restatement of the topic-segmentation wrapper's loss heads.

TEST INFRASTRUCTURE ONLY — imported by tests/ (and oracle/make_goldens_heads.py); nothing under spokennlp_b200/ imports it.

What it restates (emnlp2023-topic_segmentation/src/models/):
  * bert_for_ts.py:35-113      two views [B,2,S] -> encoder -> classifier dropout -> LossCalculator per view -> summed loss,
                               logits [B,2,S,C] (view 1 = copy of view 0 unless do_da_ts / do_tssp), cos-sim of view 0
  * modules/loss_calculator.py:25-73   ts loss ("lt": Linear + CE / focal; "cos": BCE on the pair similarities), + cl_loss_weight *
                               CSSL on the anchor view, + tssp_loss_weight * TSSP on the augmented view
  * modules/utils.py:116-138   EopPairCosineSimilarity: cosine of each labelled row with the NEXT labelled row of its example
                               (cyclic), padded with -100 to the longest example
  * modules/utils.py:141-182   FocalLoss / get_loss_fct (class weights [w0, 1-w0] when w0 != 0.5)
  * modules/tssp.py:17-35      Linear(H, num_tssp_labels) + CE on the rows where sent_token_mask != -100; returns
                               tssp_loss_weight * CE — and loss_calculator.py:71 multiplies by tssp_loss_weight AGAIN (kept)
  * modules/cssl.py:18-72,128-180,224-273   topic ids of the labelled rows, the "eop_matrix" and "eop_list" contrastive losses

Everything is written over the FLATTENED list of labelled rows (no per-example Python loops), which is also the shape a
device implementation wants.  Where the reference draws with `random.choice` (cssl.py:152-166, eop_list fall-backs) the
oracle takes the drawn indices as an input; `eop_list_indices` reproduces the reference's draw order from Python's
`random` so both sides can be run with the same seed.

Pinned against the reference itself: oracle/make_goldens_heads.py runs LossCalculator / CSSL / TSSP / the wrapper from
/root/reference and commits their outputs (tests/golden/ts_heads.pt); tests/test_oracle_heads.py holds this file to them.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class HeadsConfig:
    """The model arguments the heads read (arguments.py:51-114)."""
    num_labels: int = 2
    num_tssp_labels: int = 3
    do_da_ts: bool = False
    do_tssp: bool = False
    ts_loss_weight: float = 1.0
    ts_score_predictor: str = "lt"
    ts_score_predictor_cos_temp: float = 1.0
    focal_loss_gamma: float = 0.0
    weight_label_zero: float = 0.5
    cl_loss_weight: float = 0.0
    cl_temp: float = 1.0
    cl_anchor_level: str = "eop_matrix"
    cl_positive_k: int = 1
    cl_negative_k: int = 1
    tssp_loss_weight: float = 0.0


# ------------------------------------------------------------------------------------------------ labelled rows
def labelled_rows(labels: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Positions with a label (!= -100) in row-major order: (example index [n], position [n], count per example [B])."""
    keep = labels != -100
    b_idx, s_idx = keep.nonzero(as_tuple=True)
    return b_idx, s_idx, keep.sum(1)


def topic_ids(row_labels: Tensor, b_idx: Tensor) -> Tensor:
    """cssl.py:252-263 / utils.py:29-40: consecutive labelled rows share a topic id until a row labelled 0 (topic boundary
    AFTER it) or the end of an example.  id[j] = number of boundaries among rows < j."""
    n = row_labels.numel()
    if n == 0:
        return row_labels.new_zeros(0)
    last_of_example = torch.ones(n, dtype=torch.bool, device=row_labels.device)
    last_of_example[:-1] = b_idx[1:] != b_idx[:-1]
    boundary = ((row_labels == 0) | last_of_example).long()
    return torch.cumsum(boundary, 0) - boundary


# ------------------------------------------------------------------------------------------------ ts loss
def ts_loss_lt(logits: Tensor, labels: Tensor, gamma: float, weight_label_zero: float) -> Tensor:
    """utils.py:141-182.  gamma == 0: (weighted) mean CE over the labelled positions.  gamma != 0: `FocalLoss` — note what the
    class actually computes: it passes reduction='none' to CrossEntropyLoss.__init__ but then overwrites self.reduction with
    'mean' (utils.py:146-147), so the inner CE is ALREADY the (weighted) mean over labelled positions, a scalar; that scalar
    is multiplied by (1 - p_target)^gamma of EVERY position (ignored ones use class 0 as target, utils.py:162) and the
    products are averaged:  focal = mean_i (1 - p_i)^gamma  x  CE_mean."""
    C = logits.shape[-1]
    w = None
    if weight_label_zero != 0.5:
        w = torch.tensor([weight_label_zero, 1.0 - weight_label_zero], dtype=torch.float32, device=logits.device)
    flat, lab = logits.reshape(-1, C), labels.reshape(-1)
    ce = F.cross_entropy(flat, lab, weight=w, ignore_index=-100)
    if gamma == 0:
        return ce
    p = torch.softmax(flat, 1).gather(1, lab.clamp_min(0)[:, None])
    return ((1.0 - p) ** gamma * ce).mean()


def eop_pair_cos_sim(h: Tensor, labels: Tensor, temp: float) -> Tuple[Tensor, Tensor]:
    """utils.py:116-138 for temp != 0.  Returns (cos [B, max_n], labels [B, max_n]), both padded with -100."""
    if temp == 0:
        raise ValueError("temp == 0 (dot-product matrices) does not produce a [B, max_n] tensor in the reference either")
    B = h.shape[0]
    b_idx, s_idx, cnt = labelled_rows(labels)
    n, max_n = b_idx.numel(), int(cnt.max()) if B else 0
    start = torch.cumsum(cnt, 0) - cnt                                       # first flat row of each example
    rank = torch.arange(n, device=h.device) - start[b_idx]                                    # position of the row inside its example
    nxt = start[b_idx] + (rank + 1) % cnt[b_idx].clamp_min(1)                # cyclic successor
    rows = h[b_idx, s_idx]
    cos = F.cosine_similarity(rows, rows[nxt], dim=-1) / temp
    out = h.new_full((B, max_n), -100.0)
    lab = labels.new_full((B, max_n), -100)
    out[b_idx, rank] = cos
    lab[b_idx, rank] = labels[b_idx, s_idx]
    return out, lab


# ------------------------------------------------------------------------------------------------ CSSL
def eop_features(h: Tensor, extract_eop_segment_ids: Tensor, eop_index: Tensor) -> Tensor:
    """cssl.py:236-247: per-example segmented amax over positions sharing an id (slot k = id k), then the slots listed
    by the non-zero entries of `eop_index`, flattened over the batch."""
    B, S, H = h.shape
    pooled = torch.zeros_like(h).scatter_reduce(1, extract_eop_segment_ids[:, :, None].expand_as(h), h, reduce="amax", include_self=False)
    flat_slot = (eop_index + torch.arange(B, device=h.device)[:, None] * S).reshape(-1)
    return pooled.reshape(B * S, H)[flat_slot[eop_index.reshape(-1) != 0]]


def cssl_eop_matrix(feat: Tensor, seg: Tensor, temp: float) -> Tensor:
    """cssl.py:20-72.  For every column j: numerator = sum of exp(sim) over OTHER rows of j's topic, denominator = numerator
    + sum over rows of other topics; loss = mean of -log(num/den) over the columns whose ratio is not 0."""
    sim = F.cosine_similarity(feat[:, None, :], feat[None, :, :], dim=-1) / temp
    e = torch.exp(sim)
    same = seg[:, None] == seg[None, :]
    eye = torch.eye(seg.numel(), dtype=torch.bool, device=seg.device)
    num = (e * (same & ~eye)).sum(0)
    den = num + (e * ~same).sum(0)
    prob = num / den
    return -torch.log(prob[prob != 0]).mean()


def eop_list_indices(seg: Sequence[int], k_pos: int, k_neg: int, rng=random) -> Tuple[List[List[int]], List[List[int]]]:
    """Index choice of cssl.py:128-166 ("eop_list"): for row i of topic [start, end]: positives = the k_pos rows before i,
    falling back to a random row of [start, end) (or `end` when the topic has one row) once they run out of the topic;
    negatives = the k_neg rows after `end`, falling back to a random row of (end, last] (or of the first topic) past the
    last row.  Draws happen in the reference's order (row by row, positives then negatives) so a shared seed reproduces it."""
    n = len(seg)
    starts, ends = {}, {}
    for i, t in enumerate(seg):
        starts.setdefault(t, i)
        ends[t] = i
    first_topic = list(range(starts[seg[0]], starts[seg[0] + 1])) if (seg[0] + 1) in starts else []
    pos = [[] for _ in range(k_pos)]
    neg = [[] for _ in range(k_neg)]
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


def cssl_list(feat: Tensor, anchors: Tensor, pos: Sequence[Sequence[int]], neg: Sequence[Sequence[int]], temp: float) -> Tensor:
    """cssl.py:86-126: -log( sum_pos exp(sim) / sum_{pos+neg} exp(sim) ), averaged over the anchors."""
    def sims(idx_lists):
        return torch.stack([F.cosine_similarity(anchors, feat[torch.as_tensor(ix, dtype=torch.long, device=feat.device)], dim=-1) / temp for ix in idx_lists])
    ep, en = torch.exp(sims(pos)), torch.exp(sims(neg))
    num = ep.sum(0)
    return -torch.log(num / (num + en.sum(0))).mean()


def cssl_loss(h: Tensor, labels: Tensor, extract_eop_segment_ids: Tensor, eop_index: Tensor, cfg: HeadsConfig,
              indices: Optional[Tuple[Sequence[Sequence[int]], Sequence[Sequence[int]]]] = None, rng=random) -> Tensor:
    """cssl.py:224-273.  0 unless there are more than two labelled rows in at least two topics."""
    b_idx, s_idx, _ = labelled_rows(labels)
    seg = topic_ids(labels[b_idx, s_idx], b_idx)
    if seg.numel() <= 2 or int(seg[-1]) == 0:
        return h.new_zeros(())
    feat = eop_features(h, extract_eop_segment_ids, eop_index)
    if cfg.cl_anchor_level == "eop_matrix":
        return cssl_eop_matrix(feat, seg, cfg.cl_temp)
    if cfg.cl_anchor_level == "eop_list":
        pos, neg = indices if indices is not None else eop_list_indices(seg.tolist(), cfg.cl_positive_k, cfg.cl_negative_k, rng)
        return cssl_list(feat, feat, pos, neg, cfg.cl_temp)
    raise ValueError(f"cl_anchor_level {cfg.cl_anchor_level!r} is not restated (eop_matrix and eop_list are)")


# ------------------------------------------------------------------------------------------------ TSSP
def tssp_loss(h: Tensor, sent_token_mask: Tensor, sent_pair_orders: Tensor, w: Tensor, b: Tensor, cfg: HeadsConfig) -> Tensor:
    """tssp.py:17-35 (returns tssp_loss_weight * CE, as the reference does)."""
    feats = h[sent_token_mask != -100]
    logits = feats @ w.t() + b
    return cfg.tssp_loss_weight * F.cross_entropy(logits, sent_pair_orders[sent_pair_orders != -100])


# ------------------------------------------------------------------------------------------------ LossCalculator / wrapper
@dataclass
class HeadWeights:
    cls_w: Tensor
    cls_b: Tensor
    tssp_w: Optional[Tensor] = None
    tssp_b: Optional[Tensor] = None


def loss_calculator(h: Tensor, labels: Tensor, hw: HeadWeights, cfg: HeadsConfig, *, extract_eop_segment_ids=None, eop_index=None,
                    sent_token_mask=None, sent_pair_orders=None, da_example: bool = False, cssl_indices=None, rng=random):
    """loss_calculator.py:25-73.  Returns (loss, logits, cos [B, max_n])."""
    cos, cos_labels = eop_pair_cos_sim(h, labels, cfg.ts_score_predictor_cos_temp)
    if cfg.ts_score_predictor == "lt":
        logits = h @ hw.cls_w.t() + hw.cls_b
        ts = ts_loss_lt(logits, labels, cfg.focal_loss_gamma, cfg.weight_label_zero)
    elif cfg.ts_score_predictor == "cos":
        ts = F.binary_cross_entropy_with_logits(cos.reshape(-1), cos_labels.reshape(-1).float())     # padding (-100) included, as upstream
        logits = torch.sigmoid(cos)
    else:
        raise ValueError(cfg.ts_score_predictor)
    loss = cfg.ts_loss_weight * ts
    if not da_example and cfg.cl_loss_weight != 0:
        loss = loss + cfg.cl_loss_weight * cssl_loss(h, labels, extract_eop_segment_ids, eop_index, cfg, cssl_indices, rng)
    if da_example and cfg.tssp_loss_weight != 0:
        loss = loss + cfg.tssp_loss_weight * tssp_loss(h, sent_token_mask, sent_pair_orders, hw.tssp_w, hw.tssp_b, cfg)
    return loss, logits, cos


def wrapper_forward(encode: Callable[[Tensor, Tensor, Tensor], Tensor], hw: HeadWeights, cfg: HeadsConfig, input_ids: Tensor,
                    attention_mask: Tensor, token_type_ids: Tensor, labels: Tensor, extract_eop_segment_ids: Tensor, eop_index: Tensor,
                    sent_token_mask: Tensor, sent_pair_orders: Tensor, *, dropout: Callable[[Tensor], Tensor] = lambda x: x,
                    cssl_indices=None, rng=random):
    """bert_for_ts.py:35-113 with `encode(ids, mask, token_types) -> [B,S,H]` standing for `self.bert(...)[0]` (the oracle
    encoder on the CPU, the drop-in module on a GPU).  All inputs are the reference's [B, 2, S] pairs."""
    h0 = dropout(encode(input_ids[:, 0], attention_mask[:, 0], token_type_ids[:, 0]))
    two = cfg.do_da_ts or cfg.do_tssp
    loss, logits0, cos = loss_calculator(h0, labels[:, 0], hw, cfg, extract_eop_segment_ids=extract_eop_segment_ids[:, 0],
                                         eop_index=eop_index[:, 0], cssl_indices=cssl_indices, rng=rng)
    logits1 = logits0
    if two:
        h1 = dropout(encode(input_ids[:, 1], attention_mask[:, 1], token_type_ids[:, 1]))
        da_loss, logits1, _ = loss_calculator(h1, labels[:, 1], hw, cfg, sent_token_mask=sent_token_mask[:, 1],
                                              sent_pair_orders=sent_pair_orders[:, 1], da_example=True)
        loss = loss + da_loss
    return loss, torch.stack([logits0, logits1], 1), cos


# ------------------------------------------------------------------------------------------------ synthetic batch (Appendix A.1)
def synth_pair_batch(B: int, S: int, vocab: int, seed: int, bos_every: Tuple[int, int] = (5, 12)):
    """A [B,2,S] batch with the reference's field semantics: labelled [BOS] rows, their running
    numbering, the gather list, and for the augmented view the sentence-order classes."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(5, vocab, (B, 2, S), generator=g)
    ids[:, :, 0] = 1
    mask = torch.ones(B, 2, S, dtype=torch.long)
    tt = torch.zeros(B, 2, S, dtype=torch.long)
    labels = torch.full((B, 2, S), -100, dtype=torch.long)
    seg = torch.zeros(B, 2, S, dtype=torch.long)
    gather = torch.zeros(B, 2, S, dtype=torch.long)
    stm = torch.full((B, 2, S), -100, dtype=torch.long)
    spo = torch.full((B, 2, S), -100, dtype=torch.long)
    for b in range(B):
        for v in range(2):
            n = S if (b == 0 or v == 1) else int(torch.randint(S // 2, S + 1, (1,), generator=g))
            mask[b, v, n:] = 0
            ids[b, v, n:] = 0
            pos, p = [], 1
            while p < n - 1:
                pos.append(p)
                p += int(torch.randint(bos_every[0], bos_every[1] + 1, (1,), generator=g))
            lab = (torch.rand(len(pos), generator=g) < 0.7).long()
            keep = pos[:-1]                                   # the last sentence of a window carries no label (:845-849)
            labels[b, v, keep] = lab[:-1]
            seg[b, v, keep] = torch.arange(1, len(keep) + 1)
            gather[b, v, :len(keep) + 1] = torch.arange(0, len(keep) + 1)
            stm[b, v, pos] = lab
            spo[b, v, pos] = torch.randint(0, 3, (len(pos),), generator=g)
    return dict(input_ids=ids, attention_mask=mask, token_type_ids=tt, labels=labels, extract_eop_segment_ids=seg,
                eop_index_for_aggregate_batch_eop_features=gather, sent_token_mask=stm, sent_pair_orders=spo)
