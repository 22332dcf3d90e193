"""Pin the CPU oracle (oracle/bert_oracle.py) against outputs of the real
reference minted by oracle/make_goldens.py (HF BertModel eager + the reference's
own wrappers, run in the build container).  fp32 vs fp32: tolerance 2e-5 relative
(summation-order differences only)."""
import os

import torch

from conftest import GOLDEN, rel_err
from oracle import bert_oracle as O

TOL = 2e-5


def _load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def test_tiny_forward_matches_reference():
    g = _load("tiny_bert.pt")
    cfg = O.OracleConfig(**g["config"])
    out = O.bert_model(g["state_dict"], cfg, g["input_ids"], g["attention_mask"], g["token_type_ids"])
    assert rel_err(out.last_hidden_state, g["last_hidden_state"]) < TOL
    assert rel_err(out.pooler_output, g["pooler_output"]) < TOL
    assert len(out.hidden_states) == len(g["hidden_states"]) == cfg.num_hidden_layers + 1
    for a, b in zip(out.hidden_states, g["hidden_states"]):
        assert rel_err(a, b) < TOL
    assert len(out.attentions) == cfg.num_hidden_layers
    for a, b in zip(out.attentions, g["attentions"]):
        assert a.shape == b.shape
        assert rel_err(a, b) < TOL


def test_tiny_loss_logits_argmax_match_reference_wrapper():
    g = _load("tiny_bert.pt")
    cfg = O.OracleConfig(**g["config"])
    loss, logits = O.topicseg_loss(g["state_dict"], cfg, g["cls_w"], g["cls_b"], g["input_ids"],
                                   g["attention_mask"], g["token_type_ids"], g["labels"])
    # bert_for_ts.py wrapper output: (loss, logits[B,2,S,2], cos)
    assert abs(float(loss) - float(g["wrapper_loss"])) < 1e-5
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    assert rel_err(logits, g["wrapper_logits"][:, 0]) < TOL
    keep = g["labels"] != -100
    assert torch.equal(O.boundary_argmax(logits)[keep], g["wrapper_logits"][:, 0].argmax(-1)[keep])


def test_tiny_gradients_match_reference_autograd():
    g = _load("tiny_bert.pt")
    cfg = O.OracleConfig(**g["config"])
    sd = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
    w = g["cls_w"].clone().requires_grad_(True)
    b = g["cls_b"].clone().requires_grad_(True)
    loss, _ = O.topicseg_loss(sd, cfg, w, b, g["input_ids"], g["attention_mask"], g["token_type_ids"], g["labels"])
    loss.backward()
    for k, ref in g["grads"].items():
        got = {"classifier.weight": w, "classifier.bias": b}.get(k, sd.get(k))
        assert got is not None and got.grad is not None, k
        # key.bias gradients are mathematically zero (softmax is shift-invariant per query row):
        # both sides hold rounding noise there, so the bound is absolute + relative.
        err = float((got.grad.double() - ref.double()).norm())
        assert err <= 5e-5 * float(ref.double().norm()) + 1e-7, (k, err)
    # pooler gets no gradient (bert_for_ts.py:20 ignores it)
    assert sd["pooler.dense.weight"].grad is None


def _padidx_loss(sd, cfg, g, w, b, **inp):
    h = O.bert_model(sd, cfg, inp.get("input_ids"), g["attention_mask"], g["token_type_ids"], inputs_embeds=inp.get("inputs_embeds")).last_hidden_state
    logits = h @ w.t() + b
    loss = torch.nn.functional.cross_entropy(logits.view(-1, 2), g["labels"].view(-1)) + 0.1 * (h * g["attention_mask"][..., None]).pow(2).mean()
    return loss, h


def test_pad_token_row_gets_no_gradient_like_the_reference():
    """padding_idx of the word table (bert_model.py:171): id 0 at live, loss-carrying positions — HF autograd leaves row 0 of
    the word-table gradient exactly zero, and so does the restatement; all other gradients agree."""
    g = _load("tiny_bert_padidx.pt")
    cfg = O.OracleConfig(**g["config"])
    assert cfg.pad_token_id == 0
    sd = {k: v.clone().requires_grad_(True) for k, v in O.random_state_dict(cfg, seed=g["weight_seed"]).items()}
    w, b = g["cls_w"].clone().requires_grad_(True), g["cls_b"].clone().requires_grad_(True)
    loss, h = _padidx_loss(sd, cfg, g, w, b, input_ids=g["input_ids"])
    assert abs(float(loss) - float(g["loss"])) < 1e-5 and rel_err(h, g["last_hidden_state"]) < TOL
    loss.backward()
    assert float(sd["embeddings.word_embeddings.weight"].grad[0].abs().max()) == 0.0
    for k, ref in g["grads"].items():
        got = {"classifier.weight": w, "classifier.bias": b}.get(k, sd.get(k))
        err = float((got.grad.double() - ref.double()).norm())
        assert err <= 5e-5 * float(ref.double().norm()) + 1e-7, (k, err)
    # the same forward through inputs_embeds: gradient wrt the embeddings tensor, none for the word table
    sd = {k: v.clone().requires_grad_(True) for k, v in O.random_state_dict(cfg, seed=g["weight_seed"]).items()}
    emb = sd["embeddings.word_embeddings.weight"].detach()[g["input_ids"]].clone().requires_grad_(True)
    loss, h = _padidx_loss(sd, cfg, g, w, b, inputs_embeds=emb)
    assert abs(float(loss) - float(g["loss_embeds"])) < 1e-5
    loss.backward()
    assert sd["embeddings.word_embeddings.weight"].grad is None
    assert rel_err(emb.grad, g["d_inputs_embeds"]) < 5e-5
    for k in ("embeddings.position_embeddings.weight", "embeddings.token_type_embeddings.weight", "embeddings.LayerNorm.weight"):
        assert rel_err(sd[k].grad, g["grads_embeds"][k]) < 5e-5, k


def test_bert_base_forward_matches_reference():
    g = _load("bert_base_2x128.pt")
    cfg = O.OracleConfig(**g["config"])
    sd = O.random_state_dict(cfg, seed=g["weight_seed"])
    out = O.bert_model(sd, cfg, g["input_ids"], g["attention_mask"], g["token_type_ids"])
    assert rel_err(out.last_hidden_state, g["last_hidden_state"]) < TOL
    assert rel_err(out.pooler_output, g["pooler_output"]) < TOL
    assert rel_err(out.hidden_states[6][:, :16], g["hidden_state_6"]) < TOL
    diag = torch.diagonal(out.attentions[0][:, 9], dim1=1, dim2=2)
    assert rel_err(diag, g["attn_l0_h9_diag"]) < TOL
    logits = O.token_cls_logits(out.last_hidden_state, g["cls_w"], g["cls_b"])
    assert rel_err(logits, g["logits"]) < TOL
    margin = (g["logits"][..., 0] - g["logits"][..., 1]).abs()
    safe = margin > 1e-4
    assert torch.equal(logits.argmax(-1)[safe], g["logits"].argmax(-1)[safe])
    # ditto pooling plumbing (evaluation_ditto.py:125-155): shapes and finiteness
    pooled = O.ditto_pool(out, g["attention_mask"], layer=0, head=9)
    assert pooled.shape == (2, cfg.hidden_size) and torch.isfinite(pooled).all()


def test_mmvts_layers_match_in_tree_reference():
    g = _load("mmvts_layers.pt")
    cfg = O.OracleConfig(hidden_size=g["H"], num_attention_heads=g["heads"], intermediate_size=g["I"],
                         num_hidden_layers=1)
    add = O.additive_key_mask(g["mask01"], torch.float32, fill=-1000000.0)
    y, _ = O.bert_layer(g["self_sd"], "", cfg, g["x"], add)
    assert rel_err(y, g["y_self"]) < TOL
    yc = O.bert_cross_layer(g["cross_sd"], "", cfg, g["x"], g["kv"], add, add)
    assert rel_err(yc, g["y_cross"]) < TOL


def test_mask_fill_values_agree():
    """finfo.min (HF), -1e6 (mmvts), -1e4 (TF fork) give the same result whenever every row keeps a key."""
    g = _load("tiny_bert.pt")
    cfg = O.OracleConfig(**g["config"])
    args = (g["state_dict"], cfg, g["input_ids"], g["attention_mask"], g["token_type_ids"])
    a = O.bert_model(*args).last_hidden_state
    for fill in (-1e6, -1e4):
        assert rel_err(O.bert_model(*args, mask_fill=fill).last_hidden_state, a) < 1e-6


def test_dropout_sites_match_hf_training_mode(monkeypatch):
    """Pins WHERE the oracle applies its explicit dropout multipliers: HF BertModel (eager) is run in train() mode, live in
    this process, with torch.nn.functional.dropout patched to multiply by the same masks, in call order
    (embeddings; per layer: probabilities, attention output dense, FFN output dense — bert_model.py:209,338,373,451).
    Skipped when transformers' BertModel is not importable (it is in this image)."""
    import pytest
    tf = pytest.importorskip("transformers")
    from oracle import dropout_masks as DM
    g = _load("tiny_bert.pt")
    c = g["config"]
    cfg = O.OracleConfig(**c)
    B, S = g["input_ids"].shape
    H, L, heads = cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads
    masks = DM.bert_masks(1234, 0.1, 0.1, L, B, S, H, heads, head_site=False)
    order = [masks["emb"]]
    for i in range(L):
        p = f"encoder.layer.{i}."
        order += [masks[p + "attention.probs"], masks[p + "attention.out"], masks[p + "ffn_out"]]
    queue = list(order)

    def fake_dropout(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        m = queue.pop(0)
        assert m.numel() == x.numel(), (tuple(m.shape), tuple(x.shape))
        return x * m.reshape(x.shape).to(x.dtype)

    monkeypatch.setattr(torch.nn.functional, "dropout", fake_dropout)
    hf_cfg = tf.BertConfig(attn_implementation="eager", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                           **{k: v for k, v in c.items()})
    model = tf.BertModel(hf_cfg, add_pooling_layer=False)
    missing = model.load_state_dict({k: v for k, v in g["state_dict"].items() if not k.startswith("pooler.")}, strict=False)
    assert not [k for k in missing.missing_keys if "position_ids" not in k]
    model.train()
    ref = model(input_ids=g["input_ids"], attention_mask=g["attention_mask"], token_type_ids=g["token_type_ids"]).last_hidden_state
    assert not queue, f"{len(queue)} masks were never consumed"
    out = O.bert_model(g["state_dict"], cfg, g["input_ids"], g["attention_mask"], g["token_type_ids"], masks=masks)
    assert rel_err(out.last_hidden_state, ref.detach()) < TOL
    # and the masks really bite: eval-mode output differs
    plain = O.bert_model(g["state_dict"], cfg, g["input_ids"], g["attention_mask"], g["token_type_ids"])
    assert rel_err(plain.last_hidden_state, ref.detach()) > 1e-2
    # keep-rate of the hash
    keep = float((masks["emb"] > 0).float().mean())
    assert abs(keep - 0.9) < 0.02


def test_dropout_hash_statistics():
    """The counter-based mask (oracle/dropout_masks.py == ptx.cuh drop_hash) behaves like independent Bernoulli(1-p) draws:
    keep rate within 4 sigma of 1-p for several (seed, site) pairs, no correlation between the two 15-bit lanes of a hash,
    between neighbouring pairs, between rows, or between sites / steps."""
    import numpy as np
    from oracle import dropout_masks as D
    p, rows, H = 0.1, 512, 768
    n = rows * H
    sigma = (p * (1 - p) / n) ** 0.5
    ms = {}
    for seed, site in ((0, 0), (1, 0), (1234567, 8 * 5 + 1), (2 ** 31 - 1, 0xE000)):
        m = (D.hidden_mask(seed, site, p, rows, H).numpy() > 0).astype(np.float64)
        ms[(seed, site)] = m
        assert abs(m.mean() - (1 - p)) < 4 * sigma, (seed, site, m.mean())
        c = m - m.mean()
        var = (c * c).mean()
        for name, a, b in (("lanes of one hash", c[:, 0::2], c[:, 1::2]), ("neighbouring pairs", c[:, :-2], c[:, 2:]),
                           ("rows", c[:-1], c[1:])):
            corr = (a * b).mean() / var
            assert abs(corr) < 4.0 / (a.size ** 0.5), (seed, site, name, corr)
    a, b = ms[(0, 0)] - ms[(0, 0)].mean(), ms[(1, 0)] - ms[(1, 0)].mean()
    assert abs((a * b).mean() / (a * a).mean()) < 4.0 / (n ** 0.5)             # consecutive step seeds are unrelated
    # attention-probability site: same checks along the key axis of one (batch, head)
    pm = (D.prob_mask(77, 3 * 8, p, 1, 2, 256, 512).numpy() > 0).astype(np.float64)
    assert abs(pm.mean() - (1 - p)) < 4 * (p * (1 - p) / pm.size) ** 0.5
    c = pm - pm.mean()
    assert abs((c[..., :-1] * c[..., 1:]).mean() / (c * c).mean()) < 4.0 / (c.size ** 0.5)
    # the threshold is exact to 2^-15: p = 0.1 -> 3277 / 32768
    assert int(np.float32(0.1) * np.float32(32768.0) + np.float32(0.5)) == 3277
