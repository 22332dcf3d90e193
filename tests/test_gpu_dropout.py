"""Dropout on the GPU path (training mode, bert_model.py:209,338,373,451 and bert_for_ts.py:66-67).

The reference's masks come from torch's RNG stream and are not part of the algorithm; the B200 path draws them from a
stateless hash (oracle/dropout_masks.py restates it on the host).  Parity is therefore checked in two steps:
 1. each kernel's mask is bit-identical to the host restatement of the hash (so the masks are KNOWN), and the backward
    regenerates the forward's mask;
 2. with those known masks handed to the CPU oracle (whose dropout placement is pinned against HF train() mode in
    tests/test_oracle.py), the loss and every parameter gradient of a training step match to the same tolerances as the
    dropout-free step.
"""
import os

import pytest
import torch

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cuda.matmul.allow_tf32 = False
    from spokennlp_b200 import ops
    return ops


def _seed(v):
    return torch.tensor([v], dtype=torch.int32, device="cuda")


@pytest.mark.parametrize("rows,N,p", [(256, 768, 0.1), (300, 128, 0.5), (1024, 3072, 0.1)])
def test_gemm_residual_epilogue_mask_is_the_hash(rows, N, p):
    """out = dropout(A W^T + bias) + aux with A = 0, bias = 1, aux = 0 exposes the multipliers themselves."""
    ops = _cuda()
    from oracle import dropout_masks as DM
    K = 64
    a = torch.zeros(rows, K, dtype=torch.float16, device="cuda")
    w = torch.zeros(N, K, dtype=torch.float16, device="cuda")
    bias = torch.ones(N, dtype=torch.float32, device="cuda")
    aux = torch.zeros(rows, N, dtype=torch.float32, device="cuda")
    out = torch.empty(rows, N, dtype=torch.float32, device="cuda")
    ops.gemm(a, w, out, epilogue=ops.EPI_BIAS_RES32, bias=bias, aux=aux, drop=ops.Dropout(_seed(4242), 17, p))
    ref = DM.hidden_mask(4242, 17, p, rows, N)
    assert torch.equal(out.cpu(), ref)
    keep = float((ref > 0).float().mean())
    assert abs(keep - (1 - p)) < 0.01
    # a different site or seed gives a different mask; p = 0 gives none
    out2 = torch.empty_like(out)
    ops.gemm(a, w, out2, epilogue=ops.EPI_BIAS_RES32, bias=bias, aux=aux, drop=ops.Dropout(_seed(4242), 18, p))
    assert not torch.equal(out2, out)
    ops.gemm(a, w, out2, epilogue=ops.EPI_BIAS_RES32, bias=bias, aux=aux, drop=ops.Dropout(_seed(4243), 17, p))
    assert not torch.equal(out2, out)
    ops.gemm(a, w, out2, epilogue=ops.EPI_BIAS_RES32, bias=bias, aux=aux, drop=None)
    assert torch.equal(out2, torch.ones_like(out2))


@pytest.mark.parametrize("B,heads,Sq,Sk", [(2, 2, 128, 64), (1, 3, 200, 37), (2, 1, 64, 50)])
def test_attention_probability_mask_is_the_hash(B, heads, Sq, Sk):
    """q = k = 0 makes P uniform (1/Sk); V = one-hot(key) makes ctx[q, key] = mask[q, key] / Sk."""
    ops = _cuda()
    from oracle import dropout_masks as DM
    H, p = heads * 64, 0.1
    q = torch.zeros(B * Sq, H, dtype=torch.float16, device="cuda")
    kv = torch.zeros(B * Sk, 2 * H, dtype=torch.float16, device="cuda")
    eye = torch.eye(64, dtype=torch.float16, device="cuda")[:Sk]                  # [Sk, 64]
    for h in range(heads):
        kv.view(B, Sk, 2 * H)[:, :, H + h * 64:H + (h + 1) * 64] = eye
    ctx = torch.empty(B * Sq, H, dtype=torch.float16, device="cuda")
    lse2 = torch.empty(B, heads, Sq, dtype=torch.float32, device="cuda")
    ops.attn_fwd(q, kv, ctx, B, heads, Sq, Sk, q_col0=0, k_col0=0, v_col0=H, lse2=lse2, drop=ops.Dropout(_seed(99), 8, p))
    got = ctx.view(B, Sq, heads, 64).permute(0, 2, 1, 3)[..., :Sk].float().cpu() * Sk          # [B,h,Sq,Sk]
    ref = DM.prob_mask(99, 8, p, B, heads, Sq, Sk)
    assert torch.equal(got > 0.5, ref > 0), float(((got > 0.5) != (ref > 0)).float().mean())
    assert float((got - ref).abs().max()) < 2e-3          # fp16 rounding of P * 1/(1-p)


def test_attention_backward_regenerates_the_forward_mask():
    """attn_bwd under dropout vs autograd through `softmax -> * mask -> @ V` with the host-restated mask."""
    ops = _cuda()
    from oracle import dropout_masks as DM
    B, heads, S, p = 2, 2, 192, 0.1
    H = heads * 64
    g = torch.Generator(device="cpu").manual_seed(5)
    qkv = (torch.randn(B * S, 3 * H, generator=g) * 0.7).half().cuda()
    dctx = (torch.randn(B * S, H, generator=g) * 0.5).half().cuda()
    ctx = torch.empty(B * S, H, dtype=torch.float16, device="cuda")
    lse2 = torch.empty(B, heads, S, dtype=torch.float32, device="cuda")
    drop = ops.Dropout(_seed(777), 3, p)
    cols = dict(q_col0=0, k_col0=H, v_col0=2 * H)
    ops.attn_fwd(qkv, qkv, ctx, B, heads, S, S, lse2=lse2, drop=drop, **cols)
    dqkv = torch.empty_like(qkv)
    ws = ops.attn_bwd_workspace(B, heads, S, qkv.device)
    ops.attn_bwd(qkv, qkv, dctx, ctx, lse2, dqkv, dqkv, ws, B, heads, S, S, dq_col0=0, dk_col0=H, dv_col0=2 * H, drop=drop, **cols)
    # reference (fp32 on the same fp16-rounded operands)
    x = qkv.float().cpu().view(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)     # [3,B,h,S,d]
    mask = DM.prob_mask(777, 3, p, B, heads, S, S)
    probs = torch.softmax(x[0] @ x[1].transpose(-1, -2) / 8.0, dim=-1)
    ref_ctx = ((probs * mask) @ x[2]).permute(0, 2, 1, 3).reshape(B * S, H)
    assert rel_err(ctx.float().cpu(), ref_ctx.detach()) < 2e-3
    ref_ctx.backward(dctx.float().cpu())
    ref_d = x.grad.permute(1, 3, 0, 2, 4).reshape(B * S, 3 * H)
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        assert rel_err(dqkv[:, sl].float().cpu(), ref_d[:, sl]) < 4e-3, name


def test_layernorm_backward_emits_masked_and_plain_gradients():
    ops = _cuda()
    from oracle import dropout_masks as DM
    rows, H, p = 384, 768, 0.1
    g = torch.Generator().manual_seed(3)
    x = torch.randn(rows, H, generator=g).cuda()
    dy = (torch.randn(rows, H, generator=g) * 0.1).half().cuda()
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).cuda()
    beta = torch.zeros(H).cuda()
    mean = torch.empty(rows, device="cuda")
    rstd = torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, 1e-12, mean=mean, rstd=rstd)
    dx, dxd = (torch.empty(rows, H, dtype=torch.float16, device="cuda") for _ in range(2))
    dgam, dbet, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, dgam, dbet, dbias=dbias, dx_drop=dxd, drop=ops.Dropout(_seed(11), 5, p))
    dx0 = torch.empty_like(dx)
    dg0, db0, dbias0 = (torch.zeros(H, device="cuda") for _ in range(3))
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx0, dg0, db0, dbias=dbias0)
    assert torch.equal(dx, dx0)                                       # the residual-path gradient is untouched
    assert torch.allclose(dgam, dg0, rtol=1e-5, atol=1e-5)            # (column sums are atomics: summation order varies)
    m = DM.hidden_mask(11, 5, p, rows, H).cuda()
    assert rel_err(dxd.float(), dx0.float() * m) < 1e-3
    assert rel_err(dbias, (dx0.float() * m).sum(0)) < 2e-3


def _load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


@pytest.mark.parametrize("p_hidden,p_attn", [(0.1, 0.0), (0.0, 0.1), (0.1, 0.1)])
def test_training_step_under_dropout_matches_oracle_with_same_masks(p_hidden, p_attn):
    ops = _cuda()
    from transformers import BertConfig
    from oracle import bert_oracle as O
    from oracle import dropout_masks as DM
    from spokennlp_b200.trainer import DataParallelTrainer, TopicSegModel
    g = _load("tiny_bert.pt")
    c = g["config"]
    cfg = BertConfig(hidden_dropout_prob=p_hidden, attention_probs_dropout_prob=p_attn, **c)
    model = TopicSegModel(cfg)
    model.bert.load_state_dict({k: v for k, v in g["state_dict"].items() if not k.startswith("pooler")})
    with torch.no_grad():
        model.loss_calculator.classifier.weight.copy_(g["cls_w"])
        model.loss_calculator.classifier.bias.copy_(g["cls_b"])
    tr = DataParallelTrainer(model, lr=1e-3, total_steps=10, seed=21)
    assert tr.drop is not None
    tr._push_seed()
    seed = int(tr.seed.item())
    assert seed == tr.step_seed(0)
    batch = [g[k].cuda() for k in ("input_ids", "attention_mask", "token_type_ids", "labels")]
    tr.forward_backward(*batch)
    loss = tr.loss_value()

    B, S = g["input_ids"].shape
    ocfg = O.OracleConfig(**c)
    masks = DM.bert_masks(seed, p_hidden, p_attn, ocfg.num_hidden_layers, B, S, ocfg.hidden_size, ocfg.num_attention_heads)
    sd = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items() if not k.startswith("pooler")}
    w, b = g["cls_w"].clone().requires_grad_(True), g["cls_b"].clone().requires_grad_(True)
    ref_loss, _ = O.topicseg_loss(sd, ocfg, w, b, g["input_ids"], g["attention_mask"], g["token_type_ids"], g["labels"], masks=masks)
    ref_loss.backward()
    assert abs(loss - float(ref_loss.detach())) < 5e-4, (loss, float(ref_loss.detach()))
    assert abs(float(ref_loss) - float(g["loss"])) > 2e-5                 # dropout really changed the function
    for k, t in list(sd.items()) + [("classifier.weight", w), ("classifier.bias", b)]:
        name = "loss_calculator." + k if k.startswith("classifier") else "bert." + k
        got = tr.flat.viewg(name).double().cpu()
        ref = t.grad.double()
        err = float((got - ref).norm())
        assert err <= 1.5e-2 * float(ref.norm()) + 3e-6, (k, err, float(ref.norm()))
    # a second step draws different masks
    tr.optimizer_step()
    tr._push_seed()
    assert int(tr.seed.item()) == tr.step_seed(1) != seed


def test_graph_replay_draws_fresh_masks_and_autograd_model_trains_with_dropout():
    _cuda()
    from transformers import BertConfig
    from spokennlp_b200 import BertModel
    from spokennlp_b200.trainer import DataParallelTrainer, TopicSegModel
    g = _load("tiny_bert.pt")
    cfg = BertConfig(hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, **g["config"])
    batch = [g[k].cuda() for k in ("input_ids", "attention_mask", "token_type_ids", "labels")]
    torch.manual_seed(0)
    tr = DataParallelTrainer(TopicSegModel(cfg), lr=0.0, total_steps=20)          # lr 0: only the masks change the loss
    assert tr.capture(*batch, warmup=1)
    losses = []
    for _ in range(3):
        tr.step(*batch)
        losses.append(tr.loss_value())
    assert len({round(l, 6) for l in losses}) == 3, losses
    # autograd model: train() applies dropout (two forwards differ), eval() does not (two forwards agree bit for bit)
    m = BertModel(cfg, add_pooling_layer=False).cuda()
    m.train()
    a = m(batch[0], attention_mask=batch[1]).last_hidden_state
    b = m(batch[0], attention_mask=batch[1]).last_hidden_state
    assert not torch.equal(a, b)
    (a.float().pow(2).mean() + b.float().pow(2).mean()).backward()              # both saved mask seeds are still alive
    assert m.embeddings.word_embeddings.weight.grad.abs().sum() > 0
    m.eval()
    with torch.no_grad():
        c1 = m(batch[0], attention_mask=batch[1]).last_hidden_state
        c2 = m(batch[0], attention_mask=batch[1]).last_hidden_state
    assert torch.equal(c1, c2)
