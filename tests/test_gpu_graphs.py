"""`spokennlp_b200.graphs.GraphedStep`: a captured forward + backward of the mmvts drop-in modules replays to the eager result — on new
inputs, after the weights moved (the fp16 operand mirror is refreshed inside the graph) and with dropout (seeds advance on the device)."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def _modules(p_drop=0.0):
    from spokennlp_b200.modeling_cross import CoAttentionEncoder, LinearProjector
    H = 128
    conf = types.SimpleNamespace(hidden_size=H, num_cross_encoder_layers=2, num_cross_encoder_heads=2, intermediate_size=256,
                                 max_seq_length=512, hidden_dropout_prob=p_drop, attention_probs_dropout_prob=p_drop, ce_kv_hidden_size=2 * H,
                                 hidden_size_vis=3328, hidden_size_audio=768)
    torch.manual_seed(0)
    return LinearProjector(conf).cuda(), CoAttentionEncoder(conf).cuda(), H


def _inputs(seed, B=2, N=70, H=128):
    g = torch.Generator().manual_seed(seed)
    mask = torch.ones(B, N)
    mask[1, 50:] = 0
    return (torch.randn(B, N, H, generator=g).cuda(), torch.randn(B, N, 3328, generator=g).cuda(), torch.randn(B, N, 768, generator=g).cuda(),
            mask.cuda())


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def test_graphed_step_replays_the_eager_forward_and_backward():
    from spokennlp_b200.graphs import GraphedStep
    proj, enc, H = _modules()
    params = [p for m in (proj, enc) for p in m.parameters()]

    def step(t, v, a, mask):
        for p in params:
            p.grad = None
        pt, pv, pa = proj(t, v, a)
        ot, ov, oa = enc(mask, pt, pv, pa)
        loss = (ot * ot).mean() + (ov * ov).mean() + (oa * oa).mean()
        loss.backward()
        return loss.detach(), ot.detach()

    graphed = GraphedStep(step, _inputs(1))
    for seed in (1, 2):                                    # the captured inputs, then new ones
        x = _inputs(seed)
        loss_e, out_e = step(*x)
        grads_e = [p.grad.clone() for p in params]
        loss_g, out_g = graphed(*x)
        assert _rel(loss_g, loss_e) < 1e-5 and _rel(out_g, out_e) < 1e-5
        for p, ge in zip(params, grads_e):
            assert _rel(p.grad, ge) < 1e-4, tuple(p.shape)
    # an "optimizer step": the replay must see the new weights (fp16 mirror refreshed inside the graph)
    with torch.no_grad():
        for p in params:
            p.add_(0.01 * torch.randn_like(p))
    x = _inputs(3)
    loss_e, out_e = step(*x)
    loss_g, out_g = graphed(*x)
    assert _rel(out_g, out_e) < 1e-5 and _rel(loss_g, loss_e) < 1e-5
    torch.cuda.synchronize()


def test_graphed_step_draws_fresh_dropout_masks_on_every_replay():
    from spokennlp_b200.graphs import GraphedStep
    proj, enc, H = _modules(p_drop=0.1)
    enc.train()
    proj.eval()

    def step(t, v, a, mask):
        with torch.no_grad():
            pt, pv, pa = proj(t, v, a)
            return enc(mask, pt, pv, pa)[0]

    x = _inputs(4)
    graphed = GraphedStep(step, x)
    a = graphed(*x).clone()
    b = graphed(*x).clone()
    enc.eval()
    clean = step(*x)
    assert torch.isfinite(a).all() and torch.isfinite(b).all()
    assert _rel(a, b) > 1e-3                               # different masks
    assert 1e-3 < _rel(a, clean) < 1.0                     # dropout is on, and of the expected size
