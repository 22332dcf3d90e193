"""Mint golden vectors for the mmvts cross-modal ENCODERS and the projector from the REAL reference (build container only).

TEST INFRASTRUCTURE ONLY.   python oracle/make_goldens_mmvts.py   ->  tests/golden/mmvts_encoders.pt

What is executed (no arithmetic of ours): `MergeAttentionEncoder` (ma_encoder.py:10-71), `CoAttentionEncoder`
(ca_encoder.py:13-77) and `LinearProjector` (projector/linear_projector.py:5-30) imported from
/root/reference/mmvts/src/models behind the 3-name shim SURVEY §8c describes, in eval mode, with all three modalities
and with each single modality missing.  Weights are NOT stored: every parameter is filled from a generator seeded by its
state_dict key (`seeded_param`), so the test re-creates them by name in the drop-in modules — which at the same time holds
the drop-in's state_dict keys and shapes to the reference's."""
from __future__ import annotations

import hashlib
import importlib
import os
import sys
import types

sys.dont_write_bytecode = True

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
REF_MODELS = "/root/reference/mmvts/src/models"


def seeded_param(key: str, shape, base_seed: int = 0) -> torch.Tensor:
    """Deterministic fill of one parameter from its state_dict key: LayerNorm weights around 1, everything else N(0, 0.05)."""
    seed = int.from_bytes(hashlib.sha1(f"{base_seed}:{key}".encode()).digest()[:4], "little")
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(tuple(shape), generator=g) * 0.05
    return 1.0 + t if key.endswith("LayerNorm.weight") or ("layernorm" in key and key.endswith("weight")) else t


def fill(module, prefix_seed: int):
    with torch.no_grad():
        for k, p in module.state_dict().items():
            if p.dtype.is_floating_point:
                p.copy_(seeded_param(k, p.shape, prefix_seed))
    return module


def reference_modules():
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    for nm in ("apply_chunking_to_forward", "prune_linear_layer"):
        if not hasattr(mu, nm):
            setattr(mu, nm, getattr(pu, nm))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = lambda *a, **k: (set(), None)
    for name, sub in (("ce_pkg", "cross_encoder"), ("pj_pkg", "projector")):
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REF_MODELS, sub)]
        sys.modules[name] = pkg
    ma = importlib.import_module("ce_pkg.ma_encoder").MergeAttentionEncoder
    ca = importlib.import_module("ce_pkg.ca_encoder").CoAttentionEncoder
    lp = importlib.import_module("pj_pkg.linear_projector").LinearProjector
    return ma, ca, lp


CONF = dict(hidden_size=128, num_cross_encoder_layers=2, num_cross_encoder_heads=2, intermediate_size=256, max_seq_length=512,
            hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, ce_kv_hidden_size=256, hidden_size_vis=512, hidden_size_audio=768)


def inputs(B: int, N: int, seed: int = 41):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(B, N, CONF["hidden_size"], generator=g)
    v = torch.randn(B, N, CONF["hidden_size_vis"], generator=g)
    a = torch.randn(B, N, CONF["hidden_size_audio"], generator=g)
    mask = torch.ones(B, N)
    mask[1, 50:] = 0
    return t, v, a, mask


def main():
    MA, CA, LP = reference_modules()
    conf = types.SimpleNamespace(**CONF)
    H, B, N = CONF["hidden_size"], 2, 70
    t, v, a, mask = inputs(B, N)
    out = dict(conf=CONF, B=B, N=N)                       # inputs are regenerated from `inputs()` by the tests
    proj = fill(LP(conf).eval(), 1)
    ma = fill(MA(conf).eval(), 2)
    conf2 = types.SimpleNamespace(**dict(CONF, ce_kv_hidden_size=H))          # two-modality co-attention: K/V width = H
    ca3, ca2 = fill(CA(conf).eval(), 3), fill(CA(conf2).eval(), 4)
    out["keys"] = {nm: [(k, tuple(p.shape)) for k, p in m.state_dict().items()] for nm, m in (("proj", proj), ("ma", ma), ("ca3", ca3), ("ca2", ca2))}
    with torch.no_grad():
        pt, pv, pa = proj(t, v, a)
        out["proj"] = (pt, pv, pa)
        out["ma_tva"] = ma(mask, pt, pv, pa)
        out["ma_ta"] = tuple(x for x in ma(mask, pt, None, pa))
        out["ca_tva"] = ca3(mask, pt, pv, pa)
        out["ca_tv"] = tuple(x for x in ca2(mask, pt, pv, None))
    torch.save(out, os.path.join(OUT, "mmvts_encoders.pt"))
    for k in ("ma_tva", "ca_tva"):
        print(k, [float(x.abs().mean()) for x in out[k]])


if __name__ == "__main__":
    main()
