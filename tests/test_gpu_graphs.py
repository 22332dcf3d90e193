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


def test_graphed_step_of_the_bert_and_ponet_dropins():
    """The drop-in BertModel and PoNetModel are capture-safe too: a graph-replayed fine-tuning step (dropout on) draws new masks per
    replay and, with dropout off, reproduces the eager loss and gradients."""
    from transformers import BertConfig
    from spokennlp_b200.graphs import GraphedStep
    from spokennlp_b200.modeling_bert import BertModel
    from spokennlp_b200.modeling_ponet import PoNetConfig, PoNetModel
    kw = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, num_hidden_layers=2, vocab_size=200,
              max_position_embeddings=512, type_vocab_size=2)
    B, S = 2, 256
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(5, 200, (B, S), generator=g).cuda()
    mask = torch.ones(B, S, dtype=torch.long)
    mask[1, 200:] = 0
    mask = mask.cuda()
    seg = (torch.arange(S)[None, :] // 16 + 1).expand(B, S).contiguous().cuda()
    w = torch.randn(B, S, 128, generator=g).cuda()
    for name in ("bert", "ponet"):
        for p_drop in (0.0, 0.1):
            torch.manual_seed(1)
            if name == "bert":
                m = BertModel(BertConfig(hidden_dropout_prob=p_drop, attention_probs_dropout_prob=p_drop, **kw), add_pooling_layer=False).cuda().train()
                fwd = lambda i, a: m(i, attention_mask=a)[0]
            else:
                m = PoNetModel(PoNetConfig(hidden_dropout_prob=p_drop, attention_probs_dropout_prob=p_drop, **kw), add_pooling_layer=False).cuda().train()
                fwd = lambda i, a: m(i, attention_mask=a, segment_ids=seg)[0]
            params = [p for p in m.parameters() if p.requires_grad]

            def step(i, a):
                for p in params:
                    p.grad = None
                loss = (fwd(i, a) * w).pow(2).mean()
                loss.backward()
                return loss.detach()

            graphed = GraphedStep(step, (ids, mask))
            l1 = float(graphed(ids, mask))
            g1 = [p.grad.clone() for p in params if p.grad is not None]
            l2 = float(graphed(ids, mask))
            assert all(torch.isfinite(x).all() for x in g1), name
            if p_drop == 0.0:
                le = float(step(ids, mask))
                ge = [p.grad.clone() for p in params if p.grad is not None]
                assert abs(l1 - le) <= 1e-5 * abs(le) and abs(l2 - le) <= 1e-5 * abs(le), (name, l1, l2, le)
                assert len(g1) == len(ge) and all(_rel(a_, b_) < 1e-4 for a_, b_ in zip(g1, ge) if float(b_.norm()) > 0), name
            else:
                assert abs(l1 - l2) > 1e-6 * abs(l1), (name, l1, l2)          # fresh masks on every replay
    torch.cuda.synchronize()
