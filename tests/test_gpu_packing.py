"""SURVEY.md §8f rank 2 on the GPU: packed (variable-length) rows.  The reference right-pads every window to max_seq_length
(ts_sentence_seq_labeling.py:862-873); the packed path keeps only the valid tokens and must give, at those tokens, what the
padded path gives: attention context / LSE / gradients, the encoder output, the training loss and every parameter gradient.
Ragged cases on purpose: a full row, rows that end inside a 128-key block / a 32-row store patch / right after one token."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from spokennlp_b200 import ops
    return ops


def _mask(lens, S):
    return (torch.arange(S)[None, :] < torch.tensor(lens)[:, None]).long().cuda()


@pytest.mark.parametrize("lens,S,heads,p", [([512, 300, 129, 1], 512, 2, 0.0), ([256, 255, 33, 200, 97], 256, 4, 0.0), ([384, 17, 384], 384, 12, 0.1)])
def test_packed_attention_matches_the_padded_kernels(lens, S, heads, p):
    ops = _ops()
    B, H = len(lens), heads * 64
    g = torch.Generator().manual_seed(3)
    qkv = (torch.randn(B * S, 3 * H, generator=g) * 0.7).half().cuda()
    dctx = (torch.randn(B * S, H, generator=g) * 0.3).half().cuda()
    mask = _mask(lens, S)
    # padded QUERY rows are real rows of the padded run (they attend to the valid keys); they only contribute to dK / dV if they
    # carry gradient, which they do not in the model (no loss on padding, nobody reads their keys) — and must not here
    dctx = torch.where(mask.bool().view(-1, 1), dctx, torch.zeros_like(dctx))
    key_bias, kv_len = ops.mask_to_bias(mask)
    seed = torch.tensor([77], dtype=torch.int32, device="cuda")
    drop = ops.Dropout(seed, 3, p) if p > 0 else None
    cols = dict(q_col0=0, k_col0=H, v_col0=2 * H)
    # padded reference run
    ctx = torch.zeros(B * S, H, dtype=torch.float16, device="cuda")
    lse = torch.zeros(B, heads, S, device="cuda")
    ops.attn_fwd(qkv, qkv, ctx, B, heads, S, S, key_bias=key_bias, kv_len=kv_len, lse2=lse, drop=drop, **cols)
    dqkv = torch.zeros_like(qkv)
    ws = ops.attn_bwd_workspace(B, heads, S, "cuda")
    ops.attn_bwd(qkv, qkv, dctx, ctx, lse, dqkv, dqkv, ws, B, heads, S, S, dq_col0=0, dk_col0=H, dv_col0=2 * H, key_bias=key_bias, kv_len=kv_len,
                 drop=drop, **cols)
    # packed run on the valid rows only
    rows = ops.compact_rows(mask, 0)
    assert rows.n == sum(lens) and rows.max_n == max(lens) and rows.start.tolist() == [0] + torch.tensor(lens).cumsum(0).tolist()
    sel = rows.idx.long()
    qkv_p, dctx_p = qkv[sel].contiguous(), dctx[sel].contiguous()
    ctx_p = torch.full((rows.n, H), float("nan"), dtype=torch.float16, device="cuda")
    lse_p = torch.zeros(B, heads, S, device="cuda")
    ops.attn_fwd(qkv_p, qkv_p, ctx_p, B, heads, S, S, lse2=lse_p, drop=drop, pack=rows, **cols)
    assert torch.isfinite(ctx_p).all()                                   # every packed row was written, none twice with garbage
    assert torch.equal(ctx_p, ctx[sel])                                  # same blocks of keys in the same order: identical bits
    valid = mask.bool()[:, None, :].expand(B, heads, S)
    assert torch.equal(lse_p[valid], lse[valid])
    dqkv_p = torch.full_like(qkv_p, float("nan"))
    ws_p = ops.attn_bwd_workspace(B, heads, S, "cuda", rows=rows.n)
    ops.attn_bwd(qkv_p, qkv_p, dctx_p, ctx_p, lse_p, dqkv_p, dqkv_p, ws_p, B, heads, S, S, dq_col0=0, dk_col0=H, dv_col0=2 * H, drop=drop, pack=rows, **cols)
    assert torch.isfinite(dqkv_p).all()
    ref = dqkv[sel]
    assert torch.equal(dqkv_p[:, H:], ref[:, H:])                        # dK, dV: one writer per tile, same arithmetic
    assert rel_err(dqkv_p[:, :H].float(), ref[:, :H].float()) < 2e-3     # dQ: fp32 reduce-adds over key blocks in a run-dependent order


def test_packed_training_step_matches_the_padded_step():
    """Whole fine-tuning step (dropout 0): loss identical to the last bits of an atomic sum, parameter gradients to the tolerance
    two padded runs have between themselves; the packed run touches 45 % fewer rows."""
    _ops()
    from transformers import BertConfig
    from spokennlp_b200.trainer import DataParallelTrainer, TopicSegModel
    kw = dict(hidden_size=128, num_attention_heads=2, intermediate_size=512, num_hidden_layers=3, vocab_size=128, max_position_embeddings=256,
              type_vocab_size=2)
    lens, S = [256, 130, 64, 111], 256
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(5, 128, (4, S), generator=g)
    mask = _mask(lens, S).cpu()
    ids = torch.where(mask.bool(), ids, torch.zeros_like(ids))
    tt = torch.zeros(4, S, dtype=torch.long)
    labels = torch.full((4, S), -100, dtype=torch.long)
    labels[:, 1:250:7] = torch.randint(0, 2, (4, 36), generator=g)
    labels = torch.where(mask.bool(), labels, torch.full_like(labels, -100))
    out = []
    for pack in (False, True):
        torch.manual_seed(0)
        model = TopicSegModel(BertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **kw))
        tr = DataParallelTrainer(model, lr=1e-3, total_steps=10, seed=11)
        tr.fused_zero_grad = False
        tr.forward_backward(ids.cuda(), mask.cuda(), tt.cuda(), labels.cuda(), pack=pack)
        out.append((tr.loss_value(), tr.flat.grad32.clone()))
        tr.optimizer_step()
        torch.cuda.synchronize()
    (l0, g0), (l1, g1) = out
    assert abs(l1 - l0) <= 4e-7 * abs(l0), (l0, l1)
    assert rel_err(g1, g0) < 2e-3, rel_err(g1, g0)


def test_packed_rows_are_rejected_where_they_cannot_be_honoured():
    ops = _ops()
    from spokennlp_b200.lib import B200Error
    H = 128
    qkv = torch.zeros(64, 3 * H, dtype=torch.float16, device="cuda")
    kv = torch.zeros(64, 2 * H, dtype=torch.float16, device="cuda")
    ctx = torch.zeros(64, H, dtype=torch.float16, device="cuda")
    rows = ops.compact_rows(_mask([40, 24], 40), 0)
    with pytest.raises(B200Error):
        ops.attn_fwd(qkv, kv, ctx, 2, 2, 40, 40, q_col0=0, k_col0=0, v_col0=H, pack=rows)          # cross-attention
