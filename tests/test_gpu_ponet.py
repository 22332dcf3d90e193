"""PoNet path (BASELINE config 3) on the GPU against oracle/ponet_oracle.py.
PARITY UNPINNED: the modelscope PoNet source is absent, so the bar is "CUDA == our restatement" (1e-3 relative)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("B,S,heads,pad", [(2, 256, 2, [256, 180]), (1, 1000, 2, [777]), (3, 77, 4, [77, 1, 40]), (2, 4096, 12, [4096, 3500])])
def test_ponet_mixer_matches_restatement(B, S, heads, pad):
    _setup()
    from oracle import ponet_oracle as P
    from spokennlp_b200 import ops
    H = heads * 64
    g = torch.Generator().manual_seed(S)
    proj = torch.randn(B, S, 5 * H, generator=g).half()
    seg, mask = P.synth_segments(B, S, seed=3, pad_from=pad)
    pf = proj.float()
    ref = P.ponet_mixer(pf[..., :H], pf[..., H:2 * H], pf[..., 2 * H:3 * H], pf[..., 3 * H:4 * H], pf[..., 4 * H:], mask, seg, heads)
    kb, _ = ops.mask_to_bias(mask.cuda())
    out = torch.empty(B * S, H, dtype=torch.float16, device="cuda")
    ops.ponet_mix_fwd(proj.view(B * S, 5 * H).cuda(), seg.cuda(), out, B, S, heads, S + 2, key_bias=kb)
    assert rel_err(out.float().cpu().view(B, S, H), ref) < 1e-3
    # padded rows are exactly zero
    assert float(out.view(B, S, H)[~mask.bool().cuda()].abs().max() if (~mask.bool()).any() else 0.0) == 0.0


def test_ponet_model_forward_matches_restatement():
    _setup()
    from oracle import bert_oracle as O
    from oracle import ponet_oracle as P
    from spokennlp_b200.modeling_ponet import PoNetConfig, PoNetModel
    kw = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, num_hidden_layers=2, vocab_size=200,
              max_position_embeddings=512, type_vocab_size=2)
    sd = P.random_state_dict(O.OracleConfig(**kw), seed=5)
    m = PoNetModel(PoNetConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **kw), add_pooling_layer=False)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    m = m.cuda().eval()
    B, S = 2, 512
    ids = torch.randint(0, 200, (B, S), generator=torch.Generator().manual_seed(1))
    seg, mask = P.synth_segments(B, S, seed=7, pad_from=[512, 300])
    ref = P.ponet_model(sd, O.OracleConfig(**kw), ids, mask, None, seg)
    out = m(ids.cuda(), attention_mask=mask.cuda(), token_type_ids=None, segment_ids=seg.cuda(), output_hidden_states=True, return_dict=True)
    keep = mask.bool()
    for a, b in zip(out.hidden_states, ref):
        assert rel_err(a.cpu()[keep], b[keep]) < 1e-3
    tup = m(ids.cuda(), attention_mask=mask.cuda(), segment_ids=seg.cuda(), return_dict=False)
    # (the global branch sums the queries with float atomics: run-to-run differences are rounding-level only)
    assert torch.allclose(tup[0], out.last_hidden_state, atol=1e-4, rtol=1e-4) and tup[1] is None


@pytest.mark.parametrize("B,S,heads,pad", [(2, 512, 2, [512, 333]), (1, 1003, 2, [777]), (3, 77, 4, [77, 1, 40]), (2, 4096, 12, [4096, 3500])])
def test_ponet_mixer_backward_matches_restatement_autograd(B, S, heads, pad):
    """Sequence lengths that are no multiple of the kernels' position blocks (8 / 16 / 64 / 128), a sequence with a single kept token,
    and the BASELINE shape (12 heads: 96-thread column blocks, 4 position groups per block)."""
    _setup()
    from oracle import ponet_oracle as P
    from spokennlp_b200 import ops
    H = heads * 64
    g = torch.Generator().manual_seed(9)
    proj = torch.randn(B, S, 5 * H, generator=g).half()
    dout = (torch.randn(B, S, H, generator=g) * 0.5).half()
    seg, mask = P.synth_segments(B, S, seed=4, pad_from=pad)
    pf = proj.float().requires_grad_(True)
    ref = P.ponet_mixer(pf[..., :H], pf[..., H:2 * H], pf[..., 2 * H:3 * H], pf[..., 3 * H:4 * H], pf[..., 4 * H:], mask, seg, heads)
    ref.backward(dout.float())
    kb, _ = ops.mask_to_bias(mask.cuda())
    pc = proj.view(B * S, 5 * H).cuda()
    out = torch.empty(B * S, H, dtype=torch.float16, device="cuda")
    ws = ops.ponet_mix_fwd(pc, seg.cuda(), out, B, S, heads, S + 2, key_bias=kb)
    dproj = torch.full((B * S, 5 * H), float("nan"), dtype=torch.float16, device="cuda")
    ops.ponet_mix_bwd(pc, dout.view(B * S, H).cuda(), seg.cuda(), ws, dproj, B, S, heads, S + 2, key_bias=kb)
    got = dproj.float().cpu().view(B, S, 5 * H)
    assert torch.isfinite(got).all()
    for name, lo in (("dQ", 0), ("dK", H), ("dO", 2 * H), ("dSg", 3 * H), ("dLc", 4 * H)):
        e = rel_err(got[..., lo:lo + H], pf.grad[..., lo:lo + H])
        assert e < 3e-3, (name, e)


def test_ponet_model_gradients_match_restatement_autograd():
    _setup()
    from oracle import bert_oracle as O
    from oracle import ponet_oracle as P
    from spokennlp_b200.modeling_ponet import PoNetConfig, PoNetModel
    kw = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, num_hidden_layers=2, vocab_size=200,
              max_position_embeddings=512, type_vocab_size=2)
    sd0 = P.random_state_dict(O.OracleConfig(**kw), seed=6)
    B, S = 2, 384
    ids = torch.randint(0, 200, (B, S), generator=torch.Generator().manual_seed(2))
    seg, mask = P.synth_segments(B, S, seed=8, pad_from=[384, 250])
    w = torch.randn(B, S, 128, generator=torch.Generator().manual_seed(3)) * mask[:, :, None]
    sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    (P.ponet_model(sd, O.OracleConfig(**kw), ids, mask, None, seg)[-1] * w).sum().backward()
    m = PoNetModel(PoNetConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **kw), add_pooling_layer=False)
    m.load_state_dict(sd0)
    m = m.cuda().train()
    out = m(ids.cuda(), attention_mask=mask.cuda(), segment_ids=seg.cuda(), return_dict=True).last_hidden_state
    (out * w.cuda()).sum().backward()
    for k, p in m.named_parameters():
        ref = sd[k].grad
        err = float((p.grad.double().cpu() - ref.double()).norm())
        # The max-pooling branches are not smooth: a near-tie that the fp16 projections resolve differently from the fp32
        # restatement moves a whole gradient row to another token (and exact fp16 ties share it), so their weights get a
        # wider bound, growing with depth as the fp16 activations drift (measured 2.0e-2 at layer 0, 7.3e-2 at layer 1).
        # The kernels themselves are held to 3e-3 on identical fp16 inputs by
        # test_ponet_mixer_backward_matches_restatement_autograd; here the direction of the gradient is checked as well.
        tol = 1e-1 if ("dense_segment" in k or "dense_local" in k) else 1.5e-2
        assert err <= tol * float(ref.double().norm()) + 1e-5, (k, err, float(ref.norm()))
        if float(ref.norm()) > 1e-4:
            cos = float((p.grad.double().cpu() * ref.double()).sum() / (p.grad.double().norm() * ref.double().norm()).cpu())
            assert cos > 0.99, (k, cos)
