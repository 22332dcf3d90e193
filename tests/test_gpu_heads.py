"""SURVEY.md §8f rank 1 on the GPU: the loss heads of the topic-segmentation wrapper as device kernels (spokennlp_b200/heads.py,
csrc/heads.cuh, all through the C ABI) against golden vectors minted from the reference's own LossCalculator / CSSL / TSSP /
wrapper (tests/golden/ts_heads.pt, oracle/make_goldens_heads.py) — forward values and every gradient, in the four head
configurations the goldens hold (eop_matrix CSSL + TSSP; focal + class weights + eop_list CSSL with the reference's `random`
draws; the cosine score predictor; a ragged batch with weighted CE).  fp32 kernels against fp32 goldens: 5e-5."""
import os
import random
from types import SimpleNamespace

import pytest
import torch

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
CASES = ["full_matrix", "focal_list", "cos_only", "ragged_weighted"]
DEFAULTS = dict(num_labels=2, num_tssp_labels=3, do_da_ts=False, do_tssp=False, ts_loss_weight=1.0, ts_score_predictor="lt",
                ts_score_predictor_cos_temp=1.0, focal_loss_gamma=0.0, weight_label_zero=0.5, cl_loss_weight=0.0, cl_temp=1.0,
                cl_anchor_level="eop_matrix", cl_positive_k=1, cl_negative_k=1, tssp_loss_weight=0.0)


def _gold():
    return torch.load(os.path.join(GOLDEN, "ts_heads.pt"), weights_only=False)


def _calc(gold, rec):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from spokennlp_b200.heads import LossCalculator
    cfg = SimpleNamespace(hidden_size=gold["config"]["hidden_size"], **{**DEFAULTS, **{k: v for k, v in rec["case"].items() if k in DEFAULTS}})
    lc = LossCalculator(cfg).cuda()
    with torch.no_grad():
        lc.classifier.weight.copy_(gold["heads"]["cls_w"])
        lc.classifier.bias.copy_(gold["heads"]["cls_b"])
        lc.tssp_model.classifier.weight.copy_(gold["heads"]["tssp_w"])
        lc.tssp_model.classifier.bias.copy_(gold["heads"]["tssp_b"])
    return cfg, lc


def _close(a, b, tol=5e-5):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert rel_err(a.detach().cpu(), b) < tol, rel_err(a.detach().cpu(), b)


@pytest.mark.parametrize("name", CASES)
def test_loss_calculator_kernels_match_the_reference_modules(name):
    gold = _gold()
    rec = gold["cases"][name]
    cfg, lc = _calc(gold, rec)
    b = {k: v.cuda() for k, v in (gold["batch_ragged"] if rec.get("ragged") else gold["batch"]).items()}
    h0 = gold["h_rand"][:, 0].clone().cuda().requires_grad_(True)
    h1 = gold["h_rand"][:, 1].clone().cuda().requires_grad_(True)
    random.seed(gold["random_seed"])                       # the reference's eop_list fall-backs come from the global `random`
    l0, lg0, cs0 = lc(h0, b["labels"][:, 0], b["extract_eop_segment_ids"][:, 0], b["eop_index_for_aggregate_batch_eop_features"][:, 0])
    l1, lg1, _ = lc(h1, b["labels"][:, 1], b["extract_eop_segment_ids"][:, 1], b["eop_index_for_aggregate_batch_eop_features"][:, 1],
                    sent_token_mask=b["sent_token_mask"][:, 1], sent_pair_orders=b["sent_pair_orders"][:, 1], da_example_flag=True)
    for got, key in ((l0, "anchor_loss"), (l1, "da_loss")):
        ref = float(rec[key])
        assert abs(float(got) - ref) < 5e-5 * max(1.0, abs(ref)), (key, float(got), ref)
    _close(lg0, rec["anchor_logits"])
    _close(lg1, rec["da_logits"])
    _close(cs0, rec["anchor_cos"])
    if cfg.ts_score_predictor == "lt":                     # the boundary decision: bit-exact
        assert torch.equal(lg0.argmax(-1).cpu(), rec["anchor_logits"].argmax(-1))
    (l0 + l1).backward()
    _close(h0.grad, rec["grad_h0"])
    _close(h1.grad, rec["grad_h1"])
    for p, key in ((lc.classifier.weight, "grad_cls_w"), (lc.classifier.bias, "grad_cls_b"), (lc.tssp_model.classifier.weight, "grad_tssp_w"),
                   (lc.tssp_model.classifier.bias, "grad_tssp_b")):
        if rec[key] is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, key
        else:
            _close(p.grad, rec[key])


def test_upstream_gradient_scales_every_head():
    """loss.backward() with a non-unit upstream gradient (the wrapper sums two views; a trainer may scale the loss): the device
    scalar reaches every head's backward."""
    gold = _gold()
    rec = gold["cases"]["full_matrix"]
    _, lc = _calc(gold, rec)
    b = {k: v.cuda() for k, v in gold["batch"].items()}
    outs = []
    for scale in (1.0, 0.37):
        lc.zero_grad()
        h = gold["h_rand"][:, 0].clone().cuda().requires_grad_(True)
        loss, _, _ = lc(h, b["labels"][:, 0], b["extract_eop_segment_ids"][:, 0], b["eop_index_for_aggregate_batch_eop_features"][:, 0])
        (loss * scale).backward()
        outs.append((h.grad.clone(), lc.classifier.weight.grad.clone()))
    assert rel_err(outs[1][0], 0.37 * outs[0][0]) < 1e-5 and rel_err(outs[1][1], 0.37 * outs[0][1]) < 1e-5


def test_compaction_topic_ids_and_degenerate_inputs():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from spokennlp_b200 import heads as Hd
    # a1 a2 a3 | b1 | c1 c2 || (next example) d1 | e1 e2   (label 0 closes a topic, the end of an example closes one too)
    labels = torch.tensor([[-100, 1, 1, 0, 0, 1, 1, -100], [-100, 0, 1, 1, -100, -100, -100, -100]]).cuda()
    rows = Hd.compact(labels, -100)
    assert rows.n == 9 and rows.max_n == 6 and rows.cnt.tolist() == [6, 3] and rows.start.tolist() == [0, 6, 9]
    assert rows.idx.tolist() == [1, 2, 3, 4, 5, 6, 9, 10, 11] and rows.rank.tolist() == [0, 1, 2, 3, 4, 5, 0, 1, 2]
    lab = Hd.gather_keys(labels, rows)
    seg = torch.empty(rows.n, dtype=torch.int32, device="cuda")
    from spokennlp_b200 import lib as L
    import ctypes as C
    L.check(L.load().b200_heads_topic_ids(C.c_void_p(lab.data_ptr()), C.c_void_p(rows.ex.data_ptr()), rows.n, C.c_void_p(seg.data_ptr()), None), "topic_ids")
    assert seg.tolist() == [0, 0, 0, 1, 2, 2, 3, 4, 4]
    # long ragged lists cross the 1024-row scan chunks
    g = torch.Generator().manual_seed(0)
    big = torch.where(torch.rand(7, 700, generator=g) < 0.6, torch.randint(0, 2, (7, 700), generator=g), torch.full((7, 700), -100)).cuda()
    r2 = Hd.compact(big, -100)
    keep = (big != -100)
    assert r2.n == int(keep.sum()) and r2.idx.tolist() == keep.view(-1).nonzero().view(-1).tolist()
    lab2 = Hd.gather_keys(big, r2)
    seg2 = torch.empty(r2.n, dtype=torch.int32, device="cuda")
    L.check(L.load().b200_heads_topic_ids(C.c_void_p(lab2.data_ptr()), C.c_void_p(r2.ex.data_ptr()), r2.n, C.c_void_p(seg2.data_ptr()), None), "topic_ids")
    ex = r2.ex.cpu()
    last = torch.ones(r2.n, dtype=torch.bool)
    last[:-1] = ex[1:] != ex[:-1]
    bnd = ((lab2.cpu() == 0) | last).long()
    assert seg2.cpu().tolist() == (torch.cumsum(bnd, 0) - bnd).tolist()
    # nothing labelled: empty cosine matrix, zero contrastive term, finite loss pieces
    cfg = SimpleNamespace(hidden_size=64, **{**DEFAULTS, "cl_loss_weight": 0.5, "ts_score_predictor": "cos"})
    lc = Hd.LossCalculator(cfg).cuda()
    h = torch.randn(2, 8, 64, device="cuda", requires_grad=True)
    none = torch.full((2, 8), -100, device="cuda")
    z = torch.zeros(2, 8, dtype=torch.long, device="cuda")
    loss, logits, cos = lc(h, none, z, z)
    assert cos.shape == (2, 0) and logits.shape == (2, 0)
    with pytest.raises(Exception):
        lc(h.cpu(), none.cpu(), z.cpu(), z.cpu())            # no CPU path
