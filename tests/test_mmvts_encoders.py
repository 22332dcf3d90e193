"""mmvts cross-modal encoders and projector against outputs of the reference's own
`MergeAttentionEncoder`, `CoAttentionEncoder`, `LinearProjector` (oracle/make_goldens_mmvts.py -> tests/golden/mmvts_encoders.pt).
Weights are re-created by state_dict key on both sides, so loading them also holds the drop-in modules' parameter names
and shapes to the reference's.
  * CPU: state_dict compatibility of the drop-in modules; the oracle's composition (concatenation / chunk order, K/V pairing
    of the co-attention stacks, projector) against the reference outputs.
  * GPU: the drop-in modules themselves."""
import os
import types

import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import bert_oracle as O
from oracle.make_goldens_mmvts import inputs, seeded_param


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLDEN, "mmvts_encoders.pt"), weights_only=False)


def _conf(gold, **over):
    return types.SimpleNamespace(**dict(gold["conf"], **over))


def _sd(keys, seed):
    return {k: seeded_param(k, shape, seed) for k, shape in keys}


def _dropin(gold):
    from spokennlp_b200.modeling_cross import CoAttentionEncoder, LinearProjector, MergeAttentionEncoder
    H = gold["conf"]["hidden_size"]
    return dict(proj=(LinearProjector(_conf(gold)), 1), ma=(MergeAttentionEncoder(_conf(gold)), 2),
                ca3=(CoAttentionEncoder(_conf(gold)), 3), ca2=(CoAttentionEncoder(_conf(gold, ce_kv_hidden_size=H)), 4))


def test_dropin_modules_have_the_reference_state_dict():
    gold = torch.load(os.path.join(GOLDEN, "mmvts_encoders.pt"), weights_only=False)
    for name, (mod, seed) in _dropin(gold).items():
        mine = [(k, tuple(p.shape)) for k, p in mod.state_dict().items()]
        assert mine == gold["keys"][name], (name, set(mine) ^ set(gold["keys"][name]))
        mod.load_state_dict(_sd(gold["keys"][name], seed))                     # strict


def _oracle_cfg(gold):
    c = gold["conf"]
    return O.OracleConfig(hidden_size=c["hidden_size"], num_attention_heads=c["num_cross_encoder_heads"],
                          intermediate_size=c["intermediate_size"], num_hidden_layers=1, layer_norm_eps=1e-12)


def _layer_sd(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def test_oracle_composition_matches_the_reference_encoders(gold):
    t, v, a, mask = inputs(gold["B"], gold["N"])
    ocfg, L = _oracle_cfg(gold), gold["conf"]["num_cross_encoder_layers"]
    psd = _sd(gold["keys"]["proj"], 1)
    proj = [O.layer_norm(O.linear(x, psd[f"proj_{nm}.weight"], psd[f"proj_{nm}.bias"]), psd[f"layernorm_{nm}.weight"],
                         psd[f"layernorm_{nm}.bias"], 1e-5) for x, nm in ((t, "text"), (v, "vis"), (a, "audio"))]
    for got, ref in zip(proj, gold["proj"]):
        assert rel_err(got, ref) < 2e-5
    pt, pv, pa = gold["proj"]
    # merge attention: concatenate along the sequence, self-attention layers, split again (ma_encoder.py:40-71)
    msd = _sd(gold["keys"]["ma"], 2)
    for feats, key in (((pt, pv, pa), "ma_tva"), ((pt, pa), "ma_ta")):
        z = torch.cat(feats, 1)
        add = O.additive_key_mask(torch.cat([mask] * len(feats), 1), torch.float32, fill=-1000000.0)
        for i in range(L):
            z, _ = O.bert_layer(_layer_sd(msd, f"cross_modal_layers.{i}."), "", ocfg, z, add)
        outs = [o for o in gold[key] if o is not None]
        for got, ref in zip(torch.chunk(z, len(feats), dim=1), outs):
            assert rel_err(got, ref) < 2e-5
    # co-attention: one cross-layer stack per modality, K/V = the other modalities concatenated on the hidden dim
    # in the order (a, v) for text, (a, t) for vision, (t, v) for audio (ca_encoder.py:47-77)
    add1 = O.additive_key_mask(mask, torch.float32, fill=-1000000.0)
    csd = _sd(gold["keys"]["ca3"], 3)
    ot, ov, oa = pt, pv, pa
    for i in range(L):
        lay = lambda m: _layer_sd(csd, f"cross_modal_{m}_layers.{i}.")
        av, at, tv = torch.cat((oa, ov), -1), torch.cat((oa, ot), -1), torch.cat((ot, ov), -1)
        ot, ov, oa = (O.bert_cross_layer(lay("text"), "", ocfg, ot, av, add1, add1), O.bert_cross_layer(lay("visual"), "", ocfg, ov, at, add1, add1),
                      O.bert_cross_layer(lay("audio"), "", ocfg, oa, tv, add1, add1))
    for got, ref in zip((ot, ov, oa), gold["ca_tva"]):
        assert rel_err(got, ref) < 2e-5
    csd2 = _sd(gold["keys"]["ca2"], 4)
    ot, ov = pt, pv
    for i in range(L):
        lay = lambda m: _layer_sd(csd2, f"cross_modal_{m}_layers.{i}.")
        ot, ov = O.bert_cross_layer(lay("text"), "", ocfg, ot, ov, add1, add1), O.bert_cross_layer(lay("visual"), "", ocfg, ov, ot, add1, add1)
    assert rel_err(ot, gold["ca_tv"][0]) < 2e-5 and rel_err(ov, gold["ca_tv"][1]) < 2e-5 and gold["ca_tv"][2] is None


@pytest.mark.gpu
def test_dropin_encoders_match_the_reference_encoders(gold):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    t, v, a, mask = (x.cuda() for x in inputs(gold["B"], gold["N"]))
    mods = {}
    for name, (mod, seed) in _dropin(gold).items():
        mod.load_state_dict(_sd(gold["keys"][name], seed))
        mods[name] = mod.cuda().eval()
    tol = 1e-3                                  # fp16 operands against the reference's fp32
    with torch.no_grad():
        proj = mods["proj"](t, v, a)
        for got, ref in zip(proj, gold["proj"]):
            assert rel_err(got.cpu(), ref) < tol
        pt, pv, pa = (x.cuda() for x in gold["proj"])                          # feed the reference's projections: errors do not compound
        for got, ref in zip(mods["ma"](mask, pt, pv, pa), gold["ma_tva"]):
            assert rel_err(got.cpu(), ref) < tol
        got = mods["ma"](mask, pt, None, pa)
        assert got[1] is None and rel_err(got[0].cpu(), gold["ma_ta"][0]) < tol and rel_err(got[2].cpu(), gold["ma_ta"][2]) < tol
        for got, ref in zip(mods["ca3"](mask, pt, pv, pa), gold["ca_tva"]):
            assert rel_err(got.cpu(), ref) < tol
        got = mods["ca2"](mask, pt, pv, None)
        assert got[2] is None and rel_err(got[0].cpu(), gold["ca_tv"][0]) < tol and rel_err(got[1].cpu(), gold["ca_tv"][1]) < tol
