"""Mint golden vectors from the REAL reference, run in the build container.

TEST INFRASTRUCTURE ONLY.  Run here (needs /root/reference and `transformers`):

    python oracle/make_goldens.py

Writes `tests/golden/*.pt`.  The GPU box has no /root/reference, so the vectors
are committed; nothing at test/bench time imports the reference.

What is executed (no arithmetic of ours on this side):
  * `transformers.models.bert.modeling_bert.BertModel` with
    `attn_implementation="eager"` — the class the reference imports at
    emnlp2023-topic_segmentation/src/models/bert_for_ts.py:7,
    mmvts/src/models/text_encoder/text_encoder.py:19, ditto/evaluation_ditto.py:16,65.
  * `BertWithDAForSentenceLabelingTopicSegmentation`
    (emnlp2023-topic_segmentation/src/models/bert_for_ts.py:19-113), imported from
    /root/reference, forward under no_grad (SURVEY §8c trap 1).
  * `BertSelfAttnLayer` / `BertCrossLayer`
    (mmvts/src/models/cross_encoder/bert_model.py:456-553), imported from
    /root/reference behind the 3-name shim SURVEY §8c describes.
"""
from __future__ import annotations

import os
import sys
import types

sys.dont_write_bytecode = True  # /root/reference is read-only

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

from oracle.bert_oracle import OracleConfig, random_state_dict  # noqa: E402  (weights only)


def synth_batch(B, S, vocab, seed, pad=True):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(5, vocab, (B, S), generator=g)
    ids[:, 0] = 1
    mask = torch.ones(B, S, dtype=torch.long)
    if pad:
        for b in range(B):
            n = int(torch.randint(S // 2, S + 1, (1,), generator=g))
            if b == 0:
                n = S
            mask[b, n:] = 0
            ids[b, n:] = 0
    tt = torch.zeros(B, S, dtype=torch.long)
    labels = torch.full((B, S), -100, dtype=torch.long)
    for b in range(B):
        n = int(mask[b].sum())
        pos = torch.arange(1, n, 7)
        labels[b, pos] = (torch.rand(len(pos), generator=g) < 0.85).long()
    return ids, mask, tt, labels


def hf_bert(cfg_kwargs, sd):
    from transformers import BertConfig, BertModel
    cfg = BertConfig(attn_implementation="eager", hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0, **cfg_kwargs)
    m = BertModel(cfg)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in k or "token_type_ids" in k for k in missing), missing
    return m.eval(), cfg


def golden_tiny():
    """Tiny BERT (H=128, 2 heads of 64, I=256, L=2): full forward outputs, the
    reference wrapper's loss/logits, and autograd gradients of the lt/CE loss."""
    kw = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, num_hidden_layers=2,
              vocab_size=128, max_position_embeddings=128, type_vocab_size=2)
    ocfg = OracleConfig(**kw)
    sd = random_state_dict(ocfg, seed=7)
    m, cfg = hf_bert(kw, sd)
    B, S = 3, 128
    ids, mask, tt, labels = synth_batch(B, S, kw["vocab_size"], seed=11)
    with torch.no_grad():
        o = m(ids, attention_mask=mask, token_type_ids=tt, output_hidden_states=True,
              output_attentions=True, return_dict=True)
    g = torch.Generator().manual_seed(3)
    cls_w = torch.randn(2, 128, generator=g) * 0.05
    cls_b = torch.randn(2, generator=g) * 0.05

    # gradients: HF encoder + Linear + CE, autograd on CPU
    m.train()  # dropout probs are 0
    for p in m.parameters():
        p.grad = None
    w = cls_w.clone().requires_grad_(True)
    b = cls_b.clone().requires_grad_(True)
    h = m(ids, attention_mask=mask, token_type_ids=tt, return_dict=True).last_hidden_state
    logits = h @ w.t() + b
    loss = torch.nn.CrossEntropyLoss()(logits.reshape(-1, 2), labels.reshape(-1))
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    grads["classifier.weight"], grads["classifier.bias"] = w.grad.clone(), b.grad.clone()
    m.eval()

    # the reference wrapper (forward only, no_grad)
    sys.path.insert(0, os.path.join(REF, "emnlp2023-topic_segmentation", "src"))
    from models.bert_for_ts import BertWithDAForSentenceLabelingTopicSegmentation as Wrapper
    for k, v in dict(num_labels=2, classifier_dropout=None, do_da_ts=False, do_tssp=False, do_cssl=False,
                     ts_score_predictor="lt", ts_score_predictor_cos_temp=1, focal_loss_gamma=0.0,
                     weight_label_zero=0.5, ts_loss_weight=1.0, cl_loss_weight=0.0, tssp_loss_weight=0.0,
                     cl_temp=1, cl_anchor_level="eop_matrix", cl_positive_k=1, cl_negative_k=1,
                     num_tssp_labels=3).items():
        setattr(cfg, k, v)
    wr = Wrapper(cfg).eval()
    wr.bert.load_state_dict(sd, strict=False)
    with torch.no_grad():
        wr.loss_calculator.classifier.weight.copy_(cls_w)
        wr.loss_calculator.classifier.bias.copy_(cls_b)
        st = lambda t: torch.stack([t, t], 1)
        zeros = torch.zeros(B, 2, S, dtype=torch.long)
        wout = wr(st(ids), attention_mask=st(mask), token_type_ids=st(tt), labels=st(labels),
                  extract_eop_segment_ids=zeros, eop_index_for_aggregate_batch_eop_features=zeros,
                  sent_token_mask=zeros, sent_pair_orders=zeros)
    sys.path.pop(0)

    torch.save(dict(
        config=kw, state_dict=sd, input_ids=ids, attention_mask=mask, token_type_ids=tt, labels=labels,
        cls_w=cls_w, cls_b=cls_b,
        last_hidden_state=o.last_hidden_state, pooler_output=o.pooler_output,
        hidden_states=list(o.hidden_states), attentions=list(o.attentions),
        loss=loss.detach(), logits=logits.detach(), grads=grads,
        wrapper_loss=wout[0], wrapper_logits=wout[1], wrapper_cos=wout[2],
        source="transformers %s BertModel(eager) + /root/reference bert_for_ts.py" % __import__("transformers").__version__,
    ), os.path.join(OUT, "tiny_bert.pt"))
    print("tiny_bert: loss", float(loss), "wrapper loss", float(wout[0]))


def golden_tiny_padidx():
    """Same tiny BERT: (a) token id 0 (= pad_token_id, the word table's padding_idx — bert_model.py:171) at LIVE positions
    that receive gradient, so the golden holds the reference's "pad row gets no gradient" behaviour; (b) the same forward
    driven through `inputs_embeds` with gradients wrt the embeddings tensor and the remaining embedding parameters."""
    kw = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, num_hidden_layers=2,
              vocab_size=128, max_position_embeddings=128, type_vocab_size=2)
    sd = random_state_dict(OracleConfig(**kw), seed=7)
    m, cfg = hf_bert(kw, sd)
    assert m.embeddings.word_embeddings.padding_idx == 0
    B, S = 3, 128
    ids, mask, tt, labels = synth_batch(B, S, kw["vocab_size"], seed=12)
    g = torch.Generator().manual_seed(4)
    for b in range(B):                               # pad id inside the attended range, some of them at labelled positions
        n = int(mask[b].sum())
        where = torch.randperm(n - 1, generator=g)[:6] + 1
        ids[b, where] = 0
        ids[b, 1 + 7 * b] = 0                        # positions 1, 8, 15 carry labels (synth_batch labels every 7th token)
    tt[:, 40:] = 1
    cls_w = torch.randn(2, 128, generator=g) * 0.05
    cls_b = torch.randn(2, generator=g) * 0.05
    m.train()                                        # dropout probabilities are 0

    def run(**inp):
        for p in m.parameters():
            p.grad = None
        w, b = cls_w.clone().requires_grad_(True), cls_b.clone().requires_grad_(True)
        h = m(attention_mask=mask, token_type_ids=tt, return_dict=True, **inp).last_hidden_state
        logits = h @ w.t() + b
        # every live token contributes (mean of squares) in addition to the CE on labelled rows: padded-id rows get gradient
        loss = torch.nn.CrossEntropyLoss()(logits.reshape(-1, 2), labels.reshape(-1)) + 0.1 * (h * mask[..., None]).pow(2).mean()
        loss.backward()
        grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        grads["classifier.weight"], grads["classifier.bias"] = w.grad.clone(), b.grad.clone()
        return loss.detach(), h.detach(), grads

    loss, h, grads = run(input_ids=ids)
    assert float(grads["embeddings.word_embeddings.weight"][0].abs().max()) == 0.0
    emb = sd["embeddings.word_embeddings.weight"][ids].clone().requires_grad_(True)
    loss_e, h_e, grads_e = run(inputs_embeds=emb)
    assert "embeddings.word_embeddings.weight" not in grads_e
    torch.save(dict(config=kw, weight_seed=7, input_ids=ids, attention_mask=mask, token_type_ids=tt, labels=labels,
                    cls_w=cls_w, cls_b=cls_b, loss=loss, last_hidden_state=h, grads=grads,
                    loss_embeds=loss_e, last_hidden_state_embeds=h_e, grads_embeds=grads_e, d_inputs_embeds=emb.grad.clone(),
                    source="transformers %s BertModel(eager), padding_idx=%s" % (__import__("transformers").__version__,
                                                                                 m.embeddings.word_embeddings.padding_idx)),
               os.path.join(OUT, "tiny_bert_padidx.pt"))
    print("tiny_bert_padidx: loss", float(loss), float(loss_e), "pad-row grad", float(grads["embeddings.word_embeddings.weight"][0].abs().max()))


def golden_base():
    """BERT-base sized: weights regenerated from seed (not stored); store inputs and
    the reference's outputs for a [2,128] padded batch."""
    kw = dict(hidden_size=768, num_attention_heads=12, intermediate_size=3072, num_hidden_layers=12,
              vocab_size=30523, max_position_embeddings=512, type_vocab_size=2)
    sd = random_state_dict(OracleConfig(**kw), seed=0)
    m, _ = hf_bert(kw, sd)
    ids, mask, tt, labels = synth_batch(2, 128, kw["vocab_size"], seed=21)
    with torch.no_grad():
        o = m(ids, attention_mask=mask, token_type_ids=tt, output_hidden_states=True,
              output_attentions=True, return_dict=True)
    g = torch.Generator().manual_seed(5)
    cls_w = torch.randn(2, 768, generator=g) * 0.05
    cls_b = torch.randn(2, generator=g) * 0.05
    logits = o.last_hidden_state @ cls_w.t() + cls_b
    torch.save(dict(
        config=kw, weight_seed=0, input_ids=ids, attention_mask=mask, token_type_ids=tt, labels=labels,
        cls_w=cls_w, cls_b=cls_b,
        last_hidden_state=o.last_hidden_state, pooler_output=o.pooler_output,
        hidden_state_6=o.hidden_states[6][:, :16].clone(),
        attn_l0_h9_diag=torch.diagonal(o.attentions[0][:, 9], dim1=1, dim2=2).clone(),
        logits=logits,
    ), os.path.join(OUT, "bert_base_2x128.pt"))
    print("bert_base: |h| mean", float(o.last_hidden_state.abs().mean()))


def golden_mmvts_layers():
    """In-tree BertSelfAttnLayer / BertCrossLayer (mmvts) on a tiny config."""
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    for nm in ("apply_chunking_to_forward", "prune_linear_layer"):
        if not hasattr(mu, nm):
            setattr(mu, nm, getattr(pu, nm))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = lambda *a, **k: (set(), None)
    pkg_root = os.path.join(REF, "mmvts", "src", "models", "cross_encoder")
    # import bert_model.py as a stand-alone module (its package __init__ pulls unrelated deps)
    pkg = types.ModuleType("ce_pkg")
    pkg.__path__ = [pkg_root]
    sys.modules["ce_pkg"] = pkg
    import importlib
    bm = importlib.import_module("ce_pkg.bert_model")
    from transformers import BertConfig
    H = 128
    cfg = BertConfig(hidden_size=H, num_attention_heads=2, intermediate_size=256, num_hidden_layers=1,
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    torch.manual_seed(13)
    sl = bm.BertSelfAttnLayer(cfg, None).eval()
    cl = bm.BertCrossLayer(cfg, ce_kv_hidden_size=2 * H).eval()
    for mod in (sl, cl):
        for n_, p in mod.named_parameters():
            with torch.no_grad():
                if "LayerNorm.weight" in n_:
                    p.copy_(1 + 0.05 * torch.randn_like(p))
                else:
                    p.copy_(0.05 * torch.randn_like(p))
    B, N = 2, 90
    x = torch.randn(B, N, H)
    kv = torch.randn(B, N, 2 * H)
    valid = torch.tensor([90, 57])
    m01 = (torch.arange(N)[None, :] < valid[:, None]).float()
    add_mask = ((1.0 - m01) * -1000000.0)[:, None, None, :]      # ca_encoder.py:48-49 / ma_encoder.py:55-56
    with torch.no_grad():
        y_self = sl(x, add_mask)[0]
        y_cross = cl(x, kv, add_mask, add_mask)[0]
    torch.save(dict(H=H, heads=2, I=256, x=x, kv=kv, mask01=m01,
                    self_sd={k: v.clone() for k, v in sl.state_dict().items()},
                    cross_sd={k: v.clone() for k, v in cl.state_dict().items()},
                    y_self=y_self, y_cross=y_cross), os.path.join(OUT, "mmvts_layers.pt"))
    print("mmvts layers:", float(y_self.abs().mean()), float(y_cross.abs().mean()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    only = set(sys.argv[1:])                  # e.g. `python oracle/make_goldens.py tiny_padidx` re-mints one file
    for name, fn in (("tiny", golden_tiny), ("tiny_padidx", golden_tiny_padidx), ("base", golden_base),
                     ("mmvts_layers", golden_mmvts_layers)):
        if not only or name in only:
            fn()
