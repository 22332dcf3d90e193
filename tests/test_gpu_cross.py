"""mmvts cross-modal layers on the GPU against the committed golden vectors minted from the
reference's own in-tree `BertSelfAttnLayer` / `BertCrossLayer` (tests/golden/mmvts_layers.pt) and against the CPU
oracle's autograd for gradients.  Bar: outputs within 1e-3 relative, gradients within 1e-2 relative."""
import os
import types

import pytest
import torch

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from transformers import BertConfig
    g = torch.load(os.path.join(GOLDEN, "mmvts_layers.pt"), weights_only=False)
    cfg = BertConfig(hidden_size=g["H"], num_attention_heads=g["heads"], intermediate_size=g["I"], num_hidden_layers=1,
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    return g, cfg


def test_self_and_cross_layer_forward_match_reference_golden():
    g, cfg = _setup()
    from spokennlp_b200.modeling_cross import BertCrossLayer, BertSelfAttnLayer
    add = ((1.0 - g["mask01"]) * -1000000.0)[:, None, None, :].cuda()
    sl = BertSelfAttnLayer(cfg, None)
    assert set(sl.state_dict().keys()) == set(g["self_sd"].keys())
    sl.load_state_dict(g["self_sd"])
    sl = sl.cuda()
    with torch.no_grad():
        y = sl(g["x"].cuda(), add)[0]
    assert rel_err(y.cpu(), g["y_self"]) < 1e-3
    cl = BertCrossLayer(cfg, ce_kv_hidden_size=2 * g["H"])
    assert set(cl.state_dict().keys()) == set(g["cross_sd"].keys())
    cl.load_state_dict(g["cross_sd"])
    cl = cl.cuda()
    with torch.no_grad():
        yc = cl(g["x"].cuda(), g["kv"].cuda(), add, add)[0]
    assert rel_err(yc.cpu(), g["y_cross"]) < 1e-3


def test_cross_layer_gradients_match_oracle_autograd():
    g, cfg = _setup()
    from oracle import bert_oracle as O
    from spokennlp_b200.modeling_cross import BertCrossLayer
    ocfg = O.OracleConfig(hidden_size=g["H"], num_attention_heads=g["heads"], intermediate_size=g["I"], num_hidden_layers=1)
    sd = {k: v.clone().requires_grad_(True) for k, v in g["cross_sd"].items()}
    x = g["x"].clone().requires_grad_(True)
    kv = g["kv"].clone().requires_grad_(True)
    add = O.additive_key_mask(g["mask01"], torch.float32, fill=-1000000.0)
    w = torch.randn(2, 90, g["H"], generator=torch.Generator().manual_seed(1))
    (O.bert_cross_layer(sd, "", ocfg, x, kv, add, add) * w).sum().backward()

    cl = BertCrossLayer(cfg, ce_kv_hidden_size=2 * g["H"])
    cl.load_state_dict(g["cross_sd"])
    cl = cl.cuda().train()
    xg = g["x"].cuda().requires_grad_(True)
    kvg = g["kv"].cuda().requires_grad_(True)
    (cl(xg, kvg, add.cuda(), add.cuda())[0] * w.cuda()).sum().backward()

    def check(name, got, ref):
        err = float((got.double().cpu() - ref.double()).norm())
        scale = float(ref.double().norm())
        if name.endswith("key.bias"):
            # mathematically zero (softmax is shift-invariant per query row): both sides hold rounding noise, so the bound
            # is relative to the query-bias gradient of the same block, a non-degenerate quantity of the same construction
            scale = float(sd[name.replace("key.bias", "query.bias")].grad.double().norm())
        assert err <= 1e-2 * scale + 1e-5, (name, err, scale)
    check("x", xg.grad, x.grad)
    check("kv", kvg.grad, kv.grad)
    for k, p in cl.named_parameters():
        check(k, p.grad, sd[k].grad)


def test_encoders_and_projector_plumbing():
    """MergeAttentionEncoder / CoAttentionEncoder / LinearProjector with the reference's config fields; values are checked
    against the oracle layer by layer."""
    g, _ = _setup()
    from oracle import bert_oracle as O
    from spokennlp_b200.modeling_cross import CoAttentionEncoder, LinearProjector, MergeAttentionEncoder
    H = 128
    conf = types.SimpleNamespace(hidden_size=H, num_cross_encoder_layers=2, num_cross_encoder_heads=2, intermediate_size=256,
                                 max_seq_length=512, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, ce_kv_hidden_size=2 * H,
                                 hidden_size_vis=3328, hidden_size_audio=768)
    torch.manual_seed(0)
    B, N = 2, 70
    mask = torch.ones(B, N)
    mask[1, 50:] = 0
    t, v, a = torch.randn(B, N, H), torch.randn(B, N, 3328), torch.randn(B, N, 768)
    ocfg = O.OracleConfig(hidden_size=H, num_attention_heads=2, intermediate_size=256, num_hidden_layers=1, layer_norm_eps=1e-12)

    proj = LinearProjector(conf).cuda()
    pt, pv, pa = proj(t.cuda(), v.cuda(), a.cuda())
    for got, x, nm in ((pt, t, "text"), (pv, v, "vis"), (pa, a, "audio")):
        lin, ln = getattr(proj, f"proj_{nm}"), getattr(proj, f"layernorm_{nm}")
        ref = O.layer_norm(O.linear(x, lin.weight.detach().cpu(), lin.bias.detach().cpu()), ln.weight.detach().cpu(),
                           ln.bias.detach().cpu(), ln.eps)
        assert rel_err(got.detach().cpu(), ref) < 1e-3, nm

    ma = MergeAttentionEncoder(conf).cuda()
    with torch.no_grad():
        mt, mv, mau = ma(mask.cuda(), pt.detach(), pv.detach(), pa.detach())
    z = torch.cat((pt, pv, pa), 1).detach().cpu()
    add = O.additive_key_mask(torch.cat((mask, mask, mask), 1), torch.float32, fill=-1000000.0)
    for layer in ma.cross_modal_layers:
        z, _ = O.bert_layer({k: p.detach().cpu() for k, p in layer.state_dict().items()}, "", ocfg, z, add)
    rt, rv, ra = torch.chunk(z, 3, dim=1)
    assert rel_err(mt.cpu(), rt) < 1e-3 and rel_err(mv.cpu(), rv) < 1e-3 and rel_err(mau.cpu(), ra) < 1e-3

    ca = CoAttentionEncoder(conf).cuda()
    with torch.no_grad():
        ct, cv, cau = ca(mask.cuda(), pt.detach(), pv.detach(), pa.detach())
    add1 = O.additive_key_mask(mask, torch.float32, fill=-1000000.0)
    ot, ov, oa = pt.detach().cpu(), pv.detach().cpu(), pa.detach().cpu()
    for tl, vl, al in zip(ca.cross_modal_text_layers, ca.cross_modal_visual_layers, ca.cross_modal_audio_layers):
        sdl = lambda m: {k: p.detach().cpu() for k, p in m.state_dict().items()}
        av, at, tv = torch.cat((oa, ov), -1), torch.cat((oa, ot), -1), torch.cat((ot, ov), -1)
        ot, ov, oa = (O.bert_cross_layer(sdl(tl), "", ocfg, ot, av, add1, add1), O.bert_cross_layer(sdl(vl), "", ocfg, ov, at, add1, add1),
                      O.bert_cross_layer(sdl(al), "", ocfg, oa, tv, add1, add1))
    assert rel_err(ct.cpu(), ot) < 1e-3 and rel_err(cv.cpu(), ov) < 1e-3 and rel_err(cau.cpu(), oa) < 1e-3
