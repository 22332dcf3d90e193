"""SURVEY.md §8 row a1 end to end: the topic-segmentation wrapper (two views, classifier / focal / cosine predictor, CSSL,
TSSP) on top of an encoder, against the outputs of the reference's own wrapper (tests/golden/ts_heads.pt).

The wrapper logic is the oracle's restatement (oracle/ts_heads_oracle.py, itself pinned to the reference in
tests/test_oracle_heads.py); what this file checks is the ENCODER underneath it:
  * on the GPU: `spokennlp_b200.BertModel`, i.e. the CUDA path through the C ABI, forward and — for the gradients — its
    autograd bridge, exactly what the reference wrapper would call after INTEGRATION.md's two-line patch;
  * on the CPU (not gpu-marked): HuggingFace `BertModel`, which proves the test body itself (same helper, other device)."""
import os
import random

import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import bert_oracle as O
from oracle import ts_heads_oracle as T

CASES = ["full_matrix", "focal_list", "cos_only", "ragged_weighted"]


def _gold():
    return torch.load(os.path.join(GOLDEN, "ts_heads.pt"), weights_only=False)


def _run_wrapper(model, gold, name, device, grad=False):
    rec = gold["cases"][name]
    cfg = T.HeadsConfig(**{k: v for k, v in rec["case"].items() if k in T.HeadsConfig.__dataclass_fields__})
    b = {k: v.to(device) for k, v in (gold["batch_ragged"] if rec.get("ragged") else gold["batch"]).items()}
    hw = T.HeadWeights(**{k: v.to(device).clone().requires_grad_(grad) for k, v in gold["heads"].items()})

    def encode(ids, mask, tt):           # bert_for_ts.py:55-66: positional ids, kwargs, return_dict=False, outputs[0]
        return model(ids, attention_mask=mask, head_mask=None, token_type_ids=tt, position_ids=None, inputs_embeds=None,
                     output_attentions=None, output_hidden_states=None, return_dict=False)[0]
    random.seed(gold["random_seed"])
    with torch.set_grad_enabled(grad):
        loss, logits, cos = T.wrapper_forward(encode, hw, cfg, b["input_ids"], b["attention_mask"], b["token_type_ids"], b["labels"],
                                              b["extract_eop_segment_ids"], b["eop_index_for_aggregate_batch_eop_features"],
                                              b["sent_token_mask"], b["sent_pair_orders"])
    return cfg, rec, hw, loss, logits, cos


def _labels(gold, rec):
    return (gold["batch_ragged"] if rec.get("ragged") else gold["batch"])["labels"]


def _check_forward(cfg, rec, loss, logits, cos, labels, tol):
    ref_loss = float(rec["wrapper_loss"])
    assert abs(float(loss) - ref_loss) < 5 * tol * max(1.0, abs(ref_loss)), (float(loss), ref_loss)
    assert rel_err(logits.detach().cpu(), rec["wrapper_logits"]) < 3 * tol
    assert rel_err(cos.detach().cpu(), rec["wrapper_cos"]) < 5 * tol
    if cfg.ts_score_predictor == "lt":
        # the boundary decision, at the labelled [BOS] rows the reference scores (ts_sentence_seq_labeling.py:1032,1143): bit-exact
        # wherever the reference's logit margin exceeds the logit error bound; excluded near-ties are counted
        ref = rec["wrapper_logits"]
        margin = (ref[..., 0] - ref[..., 1]).abs()
        bound = 2 * float((logits.detach().cpu() - ref).abs().max())
        scored = (labels != -100) & (margin > bound)
        assert int(scored.sum()) >= 0.9 * int((labels != -100).sum())
        assert torch.equal(logits.detach().cpu().argmax(-1)[scored], ref.argmax(-1)[scored])


def _hf_model(gold):
    from transformers import BertConfig, BertModel
    cfg = BertConfig(attn_implementation="eager", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **gold["config"])
    m = BertModel(cfg)
    m.load_state_dict(O.random_state_dict(O.OracleConfig(**gold["config"]), seed=gold["weight_seed"]), strict=False)
    return m.eval()


@pytest.mark.parametrize("name", CASES)
def test_wrapper_on_hf_encoder_cpu(name):
    gold = _gold()
    cfg, rec, _, loss, logits, cos = _run_wrapper(_hf_model(gold), gold, name, "cpu")
    _check_forward(cfg, rec, loss, logits, cos, _labels(gold, rec), 1e-5)


def _dropin(gold):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cuda.matmul.allow_tf32 = False
    from transformers import BertConfig
    from spokennlp_b200 import BertModel
    cfg = BertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **gold["config"])
    m = BertModel(cfg)
    missing, unexpected = m.load_state_dict(O.random_state_dict(O.OracleConfig(**gold["config"]), seed=gold["weight_seed"]), strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    return m.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_wrapper_on_dropin_encoder_matches_reference_wrapper(name):
    gold = _gold()
    cfg, rec, _, loss, logits, cos = _run_wrapper(_dropin(gold).eval(), gold, name, "cuda")
    _check_forward(cfg, rec, loss, logits, cos, _labels(gold, rec), 1e-3)      # north_star: hidden states within 1e-3


@pytest.mark.gpu
def test_wrapper_gradients_through_dropin_encoder_match_cpu_autograd():
    """All heads on (CSSL eop_list + TSSP + focal): d loss / d parameters through the CUDA encoder's autograd bridge against
    HF autograd on the CPU under the same wrapper restatement."""
    gold = _gold()
    name = "focal_list"
    ref_model = _hf_model(gold).train()           # dropout probabilities are 0
    _, _, hw_ref, loss_ref, _, _ = _run_wrapper(ref_model, gold, name, "cpu", grad=True)
    loss_ref.backward()
    model = _dropin(gold).train()
    _, _, hw, loss, _, _ = _run_wrapper(model, gold, name, "cuda", grad=True)
    loss.backward()
    assert abs(float(loss) - float(loss_ref)) < 5e-3 * max(1.0, abs(float(loss_ref)))
    ref_named, named = dict(ref_model.named_parameters()), dict(model.named_parameters())
    checked = 0
    for k, p in ref_named.items():
        if p.grad is None:
            continue
        got = named[k].grad
        assert got is not None, k
        err = float((got.double().cpu() - p.grad.double()).norm())
        assert err <= 2e-2 * float(p.grad.double().norm()) + 2e-6, (k, err, float(p.grad.norm()))       # fp16 activations / gradients
        checked += 1
    assert checked >= 30
    for a, r in ((hw.cls_w, hw_ref.cls_w), (hw.tssp_w, hw_ref.tssp_w)):
        assert rel_err(a.grad.cpu(), r.grad) < 2e-2
