"""The CPU restatement of the topic-segmentation loss heads (oracle/ts_heads_oracle.py) against golden vectors minted from
the reference's own LossCalculator / CSSL / TSSP / wrapper (oracle/make_goldens_heads.py -> tests/golden/ts_heads.pt).
fp32 against fp32: 2e-5 relative."""
import os
import random

import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import bert_oracle as O
from oracle import ts_heads_oracle as T


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLDEN, "ts_heads.pt"), weights_only=False)


def _cfg(case):
    fields = T.HeadsConfig.__dataclass_fields__
    return T.HeadsConfig(**{k: v for k, v in case.items() if k in fields})


def _close(a, b, tol=2e-5):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert rel_err(a, b) < tol, rel_err(a, b)


CASES = ["full_matrix", "focal_list", "cos_only", "ragged_weighted"]


@pytest.mark.parametrize("name", CASES)
def test_heads_forward_and_gradients_match_the_reference_modules(gold, name):
    rec = gold["cases"][name]
    b = gold["batch_ragged"] if rec.get("ragged") else gold["batch"]
    cfg = _cfg(rec["case"])
    hw = T.HeadWeights(**{k: v.clone().requires_grad_(True) for k, v in gold["heads"].items()})
    h0 = gold["h_rand"][:, 0].clone().requires_grad_(True)
    h1 = gold["h_rand"][:, 1].clone().requires_grad_(True)
    random.seed(gold["random_seed"])
    l0, lg0, cs0 = T.loss_calculator(h0, b["labels"][:, 0], hw, cfg, extract_eop_segment_ids=b["extract_eop_segment_ids"][:, 0],
                                     eop_index=b["eop_index_for_aggregate_batch_eop_features"][:, 0])
    l1, lg1, _ = T.loss_calculator(h1, b["labels"][:, 1], hw, cfg, sent_token_mask=b["sent_token_mask"][:, 1],
                                   sent_pair_orders=b["sent_pair_orders"][:, 1], da_example=True)
    assert abs(float(l0) - float(rec["anchor_loss"])) < 2e-5 * max(1.0, abs(float(rec["anchor_loss"])))
    assert abs(float(l1) - float(rec["da_loss"])) < 2e-5 * max(1.0, abs(float(rec["da_loss"])))
    _close(lg0.detach(), rec["anchor_logits"])
    _close(lg1.detach(), rec["da_logits"])
    _close(cs0.detach(), rec["anchor_cos"])
    (l0 + l1).backward()
    _close(h0.grad, rec["grad_h0"], 5e-5)
    _close(h1.grad, rec["grad_h1"], 5e-5)
    for mine, key in ((hw.cls_w, "grad_cls_w"), (hw.cls_b, "grad_cls_b"), (hw.tssp_w, "grad_tssp_w"), (hw.tssp_b, "grad_tssp_b")):
        if rec[key] is None:
            assert mine.grad is None or float(mine.grad.abs().max()) == 0.0
        else:
            _close(mine.grad, rec[key], 5e-5)


@pytest.mark.parametrize("name", CASES)
def test_wrapper_forward_matches_the_reference_wrapper(gold, name):
    rec = gold["cases"][name]
    b = gold["batch_ragged"] if rec.get("ragged") else gold["batch"]
    cfg = _cfg(rec["case"])
    ocfg = O.OracleConfig(**gold["config"])
    sd = O.random_state_dict(ocfg, seed=gold["weight_seed"])
    hw = T.HeadWeights(**gold["heads"])

    def encode(ids, mask, tt):
        return O.bert_model(sd, ocfg, ids, mask, tt).last_hidden_state
    random.seed(gold["random_seed"])
    with torch.no_grad():
        loss, logits, cos = T.wrapper_forward(encode, hw, cfg, b["input_ids"], b["attention_mask"], b["token_type_ids"], b["labels"],
                                              b["extract_eop_segment_ids"], b["eop_index_for_aggregate_batch_eop_features"],
                                              b["sent_token_mask"], b["sent_pair_orders"])
    assert abs(float(loss) - float(rec["wrapper_loss"])) < 2e-5 * max(1.0, abs(float(rec["wrapper_loss"])))
    _close(logits, rec["wrapper_logits"])
    _close(cos, rec["wrapper_cos"])
    if cfg.ts_score_predictor == "lt":                                   # the boundary decision itself: bit-exact
        assert torch.equal(logits.argmax(-1), rec["wrapper_logits"].argmax(-1))


def test_topic_ids_and_list_indices_follow_the_reference_rule():
    # a1 a2 a3 | b1 | c1 c2 || (next example) d1 | e1 e2      (label 0 closes a topic, the end of an example closes one too)
    labels = torch.tensor([[-100, 1, 1, 0, 0, 1, 1, -100], [-100, 0, 1, 1, -100, -100, -100, -100]])
    b_idx, s_idx, cnt = T.labelled_rows(labels)
    assert cnt.tolist() == [6, 3]
    seg = T.topic_ids(labels[b_idx, s_idx], b_idx)
    assert seg.tolist() == [0, 0, 0, 1, 2, 2, 3, 4, 4]
    pos, neg = T.eop_list_indices(seg.tolist(), 1, 2, rng=random.Random(0))
    # positives: the previous row of the same topic where there is one
    assert [pos[0][i] for i in (1, 2, 5, 8)] == [0, 1, 4, 7]
    assert pos[0][3] == 3 and pos[0][6] == 6                               # single-row topics fall back to themselves
    # negatives: the rows right after the topic's last row
    assert [neg[0][i] for i in (0, 1, 2)] == [3, 3, 3] and [neg[1][i] for i in (0, 1, 2)] == [4, 4, 4]
    assert all(j in (0, 1, 2) for j in (neg[0][7], neg[0][8]))             # past the end: a row of the first topic


def test_empty_and_degenerate_inputs():
    cfg = T.HeadsConfig(cl_loss_weight=0.5)
    h = torch.randn(2, 8, 16)
    none = torch.full((2, 8), -100)
    cos, lab = T.eop_pair_cos_sim(h, none, 1.0)
    assert cos.shape == (2, 0) and lab.shape == (2, 0)
    z = torch.zeros(2, 8, dtype=torch.long)
    assert float(T.cssl_loss(h, none, z, z, cfg)) == 0.0                   # nothing labelled -> no contrastive term
    one_topic = none.clone()
    one_topic[0, 1:4] = 1
    assert float(T.cssl_loss(h, one_topic, z, z, cfg)) == 0.0              # a single topic -> no contrastive term
    with pytest.raises(ValueError):
        T.eop_pair_cos_sim(h, one_topic, 0.0)
