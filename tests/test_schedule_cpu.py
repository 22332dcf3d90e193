"""Host-side kernel schedule of the encoder (spokennlp_b200/engine.py + blocks.py), checked on the CPU with a recording
stand-in for `spokennlp_b200.ops`: which C-ABI entry point is enqueued, in which order, on which buffers.  No arithmetic
runs (the stand-ins only allocate), so this is a test of the orchestration: the round-1 default schedule must stay exactly
what the GPU runs validated, and the opt-in variants (blocks.Experimental) must only ever accumulate in place
into a residual buffer nobody else holds."""
import pytest
import torch

from spokennlp_b200 import blocks, engine, ops


class Recorder:
    def __init__(self):
        self.calls = []

    def names(self):
        return [c[0] for c in self.calls]


@pytest.fixture
def rec(monkeypatch):
    r = Recorder()

    def gemm(a, b, out, *, a_layout=0, b_layout=0, epilogue=ops.EPI_STORE, bias=None, aux=None, out2=None, alpha=None, k_splits=1, drop=None):
        r.calls.append((f"gemm/epi{epilogue}/a{a_layout}b{b_layout}", dict(out=out, aux=aux)))
        return out

    def gemm_resadd(a, b, out32, bias, *, drop=None):
        r.calls.append(("gemm_resadd", dict(out=out32)))
        return out32

    def gemm_dgrad_delta(dy, w, ctx, dctx, ws, B, heads, Sq):
        r.calls.append(("gemm_dgrad_delta", dict(out=dctx)))
        return dctx

    def attn_fwd(q, kv, ctx, *a, **k):
        r.calls.append(("attn_fwd", {}))
        return ctx

    def attn_bwd(*a, delta_ready=False, **k):
        r.calls.append(("attn_bwd" + ("/delta_ready" if delta_ready else ""), {}))

    def layernorm_fwd(x, gamma, beta, eps, *, y=None, y32=None, mean=None, rstd=None):
        r.calls.append(("layernorm_fwd", dict(x=x, y32=y32)))
        return torch.empty(x.shape, dtype=torch.float16)

    def layernorm_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, **k):
        r.calls.append(("layernorm_bwd", dict(x=x)))
        return dx

    def embed_ln_fwd(ids, tt, pos, inputs_embeds, word, pos_tab, type_tab, gamma, beta, eps, rows, S, H, *, y=None, y32=None, drop=None):
        r.calls.append(("embed_ln_fwd", dict(y32=y32)))
        return torch.empty(rows, H, dtype=torch.float16)

    def simple(name, ret=None):
        def f(*a, **k):
            r.calls.append((name, {}))
            return ret(*a, **k) if ret else None
        return f
    for name, fn in dict(gemm=gemm, gemm_resadd=gemm_resadd, gemm_dgrad_delta=gemm_dgrad_delta, gemm_dgelu_colsum=simple("gemm_dgelu_colsum"), attn_fwd=attn_fwd, attn_bwd=attn_bwd,
                         layernorm_fwd=layernorm_fwd, layernorm_bwd=layernorm_bwd, embed_ln_fwd=embed_ln_fwd,
                         embed_ln_bwd=simple("embed_ln_bwd"), colsum=simple("colsum"),
                         cast_f32_to_f16=simple("cast_f32_to_f16"),
                         attn_bwd_workspace=lambda B, heads, Sq, device, rows=None: torch.empty(B * heads * Sq * 65)).items():
        monkeypatch.setattr(ops, name, fn)
    yield r
    blocks.Experimental.from_env(None)          # back to the environment's / the default variant set


H, HEADS, INTER, L, B, S = 128, 2, 256, 3, 2, 64


def _engine():
    from transformers import BertConfig
    from spokennlp_b200 import BertModel
    cfg = BertConfig(hidden_size=H, num_attention_heads=HEADS, intermediate_size=INTER, num_hidden_layers=L, vocab_size=50,
                     max_position_embeddings=S)
    m = BertModel(cfg, add_pooling_layer=False)
    named = m._hot_named_params()
    flat = engine.FlatParams(named, "cpu")
    flat.ensure_grad()
    return engine.EncoderEngine(flat, H, HEADS, INTER, L, 1e-12)


def _forward(eng, **kw):
    ids = torch.zeros(B * S, dtype=torch.long)
    return eng.forward(ids, None, None, None, None, None, B, S, **kw)


FWD_LAYER = ["gemm/epi1/a0b0", "attn_fwd", "gemm/epi7/a0b0", "layernorm_fwd", "gemm/epi2/a0b0", "gemm/epi7/a0b0", "layernorm_fwd"]
BWD_LAYER = ["layernorm_bwd", "gemm/epi6/a1b1", "gemm/epi4/a0b1", "colsum", "gemm/epi6/a1b1", "gemm/epi5/a0b1",          # FFN block
             "layernorm_bwd", "gemm/epi6/a1b1", "gemm/epi0/a0b1", "attn_bwd", "colsum", "gemm/epi6/a1b1", "gemm/epi5/a0b1"]  # attention block


def test_round1_schedule_is_still_selectable(rec):
    blocks.Experimental.from_env("none")
    eng = _engine()
    rec.calls.clear()
    x16, x32, saved, hiddens, probs = _forward(eng, save=True)
    assert rec.names() == ["embed_ln_fwd"] + FWD_LAYER * L
    rec.calls.clear()
    eng.backward(saved, torch.empty(B * S, H, dtype=torch.float16), torch.ones(1))
    assert rec.names() == BWD_LAYER * L + ["embed_ln_bwd"]


@pytest.mark.parametrize("variants,name", [("resadd", "gemm_resadd")])
def test_resadd_accumulates_in_place_only_into_buffers_the_engine_owns(rec, variants, name):
    blocks.Experimental.from_env(variants)
    eng = _engine()
    rec.calls.clear()
    x16, x32, saved, hiddens, _ = _forward(eng, save=True)
    fwd = [n.replace("gemm/epi7/a0b0", name) for n in FWD_LAYER]
    assert rec.names() == ["embed_ln_fwd"] + fwd * L
    # every in-place accumulate targets the fp32 copy written by the LayerNorm (or the embeddings) right before it, and the
    # LayerNorm that follows reads that very buffer as its pre-LayerNorm sum (which is what the backward gets as `pre`)
    last_y32 = None
    for i, (n, a) in enumerate(rec.calls):
        if n in ("embed_ln_fwd", "layernorm_fwd"):
            last_y32 = a["y32"]
        if n == name:
            assert a["out"] is last_y32
            assert rec.calls[[j for j in range(i + 1, len(rec.calls)) if rec.calls[j][0] == "layernorm_fwd"][0]][1]["x"] is a["out"]
    assert all(sv.attn.pre is not None and sv.ffn.pre is not None for sv in saved.layers)
    assert x32 is last_y32                      # the model output is the last LayerNorm's buffer: nothing accumulated into it

    # with hidden states requested, a layer's input is handed to the caller and must survive: only the FFN block (whose
    # residual is the attention block's private output) may accumulate in place
    rec.calls.clear()
    x16, x32, _, hiddens, _ = _forward(eng, save=False, want_hidden=True)
    mixed = list(FWD_LAYER)
    mixed[5] = name
    assert rec.names() == ["embed_ln_fwd"] + mixed * L
    assert len(hiddens) == L + 1
    for n, a in rec.calls:
        if n == name:
            assert all(a["out"] is not h for h in hiddens)


def test_delta_variant_replaces_the_plain_dgrad_and_skips_the_row_statistic_pass(rec):
    blocks.Experimental.from_env("delta")
    eng = _engine()
    _, _, saved, _, _ = _forward(eng, save=True)
    rec.calls.clear()
    eng.backward(saved, torch.empty(B * S, H, dtype=torch.float16), torch.ones(1))
    bwd = [{"gemm/epi0/a0b1": "gemm_dgrad_delta", "attn_bwd": "attn_bwd/delta_ready"}.get(n, n) for n in BWD_LAYER]
    assert rec.names() == bwd * L + ["embed_ln_bwd"]


def test_default_schedule_is_the_round2_validated_set(rec, monkeypatch):
    """resadd + delta + colsum + dq16 were validated and measured on a B200 (profiles/r02a_*, r02c_*) and are the default."""
    monkeypatch.delenv("B200_EXP", raising=False)
    blocks.Experimental.from_env(None)
    assert blocks.Experimental.active() == ["resadd", "delta", "colsum", "dq16"]
    eng = _engine()
    rec.calls.clear()
    _, _, saved, _, _ = _forward(eng, save=True)
    fwd = [n.replace("gemm/epi7/a0b0", "gemm_resadd") for n in FWD_LAYER]
    assert rec.names() == ["embed_ln_fwd"] + fwd * L
    rec.calls.clear()
    eng.backward(saved, torch.empty(B * S, H, dtype=torch.float16), torch.ones(1))
    bwd = [{"gemm/epi0/a0b1": "gemm_dgrad_delta", "attn_bwd": "attn_bwd/delta_ready"}.get(n, n) for n in BWD_LAYER]
    assert bwd[2:4] == ["gemm/epi4/a0b1", "colsum"]
    bwd[2:4] = ["gemm_dgelu_colsum"]
    assert rec.names() == bwd * L + ["embed_ln_bwd"]


def test_colsum_variant_folds_the_ffn_bias_gradient_into_the_dgelu_dgrad(rec):
    blocks.Experimental.from_env("colsum")
    eng = _engine()
    _, _, saved, _, _ = _forward(eng, save=True)
    rec.calls.clear()
    eng.backward(saved, torch.empty(B * S, H, dtype=torch.float16), torch.ones(1))
    bwd = list(BWD_LAYER)
    assert bwd[2:4] == ["gemm/epi4/a0b1", "colsum"]
    bwd[2:4] = ["gemm_dgelu_colsum"]
    assert rec.names() == bwd * L + ["embed_ln_bwd"]


def test_unknown_variant_names_are_refused():
    with pytest.raises(ValueError):
        blocks.Experimental.from_env("resad")
    blocks.Experimental.from_env("none")
    assert not blocks.Experimental.active()
    blocks.Experimental.from_env(None)
