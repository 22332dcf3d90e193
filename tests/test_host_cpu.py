"""CPU-side checks (no GPU): the C-ABI library builds, loads and exports every symbol include/b200enc.h declares;
the drop-in module has HF-identical state_dict keys and refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from spokennlp_b200 import lib as L
    if L.is_stale():
        L.build()
    return L


def test_library_carries_the_hash_of_the_sources_it_was_built_from(lib, tmp_path, monkeypatch):
    """A library built from other sources than this tree's must be noticed (round 1 lost a GPU call to a stale .so)."""
    assert lib.built_hash() == lib.source_hash() and not lib.is_stale()
    assert lib.load().b200_source_hash().decode() == lib.source_hash()
    fake = tmp_path / "libb200enc.so"
    fake.write_bytes(b"\x7fELF....B200SRC:" + b"0" * 40 + b"\0....")
    monkeypatch.setattr(lib, "LIB_PATH", str(fake))
    assert lib.built_hash() == "0" * 40 and lib.is_stale()
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "absent.so"))
    assert lib.built_hash() is None and lib.is_stale()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "b200enc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    so = ctypes.CDLL(lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(so, name), f"{name} declared in include/b200enc.h but not exported"
    bound = set(lib.exported_symbols())
    assert declared == bound, (declared - bound, bound - declared)
    assert lib.load().b200_version() >= 100


def test_library_reports_errors_without_gpu(lib):
    so = lib.load()
    # argument validation happens before any CUDA call, so it is testable here
    rc = so.b200_gemm_f16(None, 0, 0, None, 0, 0, 0, 0, 0, 0, None, None, 0, None, 0, 0, None, 0, None, 1, None)
    assert rc == -1 and b"empty problem" in so.b200_last_error()
    with pytest.raises(lib.B200Error):
        lib.check(rc, "b200_gemm_f16")


def test_dropin_state_dict_keys_match_hf_and_no_cpu_fallback():
    from transformers import BertConfig
    from transformers import BertModel as HFBert
    from spokennlp_b200 import BertModel
    cfg = BertConfig(hidden_size=128, num_attention_heads=2, intermediate_size=256, num_hidden_layers=2, vocab_size=100,
                     max_position_embeddings=64)
    ours, hf = BertModel(cfg), HFBert(cfg)
    assert list(ours.state_dict().keys()) == list(hf.state_dict().keys())
    for (k, a), (_, b) in zip(ours.state_dict().items(), hf.state_dict().items()):
        assert a.shape == b.shape, k
    ours.load_state_dict(hf.state_dict())                      # reference checkpoints load unchanged
    nopool = BertModel(cfg, add_pooling_layer=False)
    assert not any(k.startswith("pooler") for k in nopool.state_dict())
    ours.resize_token_embeddings(101)                          # ts_sentence_seq_labeling.py:284
    assert ours.embeddings.word_embeddings.weight.shape[0] == 101
    from spokennlp_b200.lib import B200Error
    with pytest.raises(B200Error):
        ours(torch.zeros(1, 8, dtype=torch.long))
    with pytest.raises(B200Error):
        BertModel(BertConfig(hidden_size=96, num_attention_heads=2))   # head_dim != 64


def test_wgrad_split_heuristic_never_leaves_empty_splits():
    from spokennlp_b200.ops import wgrad_splits
    for Mo, Ni, T in [(768, 768, 16384), (2304, 768, 16384), (3072, 768, 16384), (768, 3072, 16384), (128, 256, 64),
                      (768, 768, 300), (2304, 768, 4096 * 2)]:
        s = wgrad_splits(Mo, Ni, T)
        kb = (T + 63) // 64
        per = (kb + s - 1) // s
        assert s >= 1 and per * (s - 1) < kb


def test_every_ops_and_library_reference_resolves():
    """Static check (no GPU): every `ops.<name>` used by the host package exists, and every library symbol `ops.py` calls
    is bound in lib.py — so a refactor cannot leave the GPU-only code paths pointing at missing functions."""
    import ast
    import glob
    from spokennlp_b200 import lib as L
    from spokennlp_b200 import ops
    pkg = os.path.join(ROOT, "spokennlp_b200")
    files = glob.glob(os.path.join(pkg, "*.py")) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + \
        glob.glob(os.path.join(ROOT, "tests", "test_gpu_*.py")) + glob.glob(os.path.join(ROOT, "tools", "*.py"))
    missing = []
    for f in files:
        tree = ast.parse(open(f).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == "ops":
                if not hasattr(ops, node.attr):
                    missing.append((os.path.basename(f), node.attr))
    assert not missing, missing
    src = open(os.path.join(pkg, "ops.py")).read()
    used = set(re.findall(r"\.(b200_[a-z0-9_]+)\(", src))
    assert used <= set(L.exported_symbols()), used - set(L.exported_symbols())


def test_oracle_is_test_infrastructure_only():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import oracle/ (the product path must
    never route through the CPU restatement)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    offenders = []
    for base in ("spokennlp_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(root, base)):
            for f in files:
                if f.endswith(".py") and pat.search(open(os.path.join(dirpath, f)).read()):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_bench_synthetic_batch_follows_the_spec():
    """ids uniform in [1000, 30522) with [CLS] = 101 first and the added [BOS] id 30522 at every
    position 1 + 20k; labels are -100 except at [BOS] positions (and on padding); disjoint seeded shards per rank."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ids, mask, tt, labels = bench.synth_batch(torch, 32, 512, 1234)
    assert ids.shape == mask.shape == tt.shape == labels.shape == (32, 512) and ids.dtype == torch.int64
    bos = torch.arange(1, 512, 20)
    assert bool((ids[:, 0] == 101).all()) and bool((ids[:, bos] == 30522).all()) and int(ids.max()) < bench.VOCAB
    other = torch.ones(512, dtype=torch.bool)
    other[0] = False
    other[bos] = False
    assert bool((ids[:, other] >= 1000).all()) and bool((ids[:, other] < 30522).all())
    assert bool(mask.all()) and not bool(tt.any())
    assert bool((labels[:, other] == -100).all()) and bool((labels[:, 0] == -100).all())
    assert set(labels[:, bos].unique().tolist()) <= {0, 1} and 0.7 < float(labels[:, bos].float().mean()) < 0.95
    ids2, _, _, _ = bench.synth_batch(torch, 32, 512, 1235)
    assert not torch.equal(ids, ids2)                                        # rank r uses seed 1234 + r
    pids, pmask, _, plabels = bench.synth_batch(torch, 8, 512, 7, padded=True)
    lens = pmask.sum(1)
    assert int(lens.min()) >= 256 and int(lens.max()) <= 512
    assert bool((plabels[pmask == 0] == -100).all())
    assert bench.FLOP_PER_SEQ == 3 * 12 * (24 * 512 * 768 * 768 + 4 * 512 * 512 * 768)   # fwd+bwd, SURVEY §8d


def test_mma_issue_sites_are_guarded_by_elect_not_by_a_lane_compare():
    """A `tcgen05.mma` sequence under `if (lane == 0)` compiles to an ELECT / broadcast loop with a branch per MMA (~90 clk
    per MMA in the attention backward, 6-25 % on the GEMM shapes); under `elect_one_sync()` it is 4-5 uniform instructions.  Hold the
    kernels to the latter: no umma call may sit within a few lines of a lane-0 guard."""
    import re
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spokennlp_b200", "csrc")
    bad, seen = [], 0
    for name in ("gemm2.cuh", "attn_fwd3.cuh", "attn_bwd3.cuh"):
        lines = open(os.path.join(csrc, name)).read().split("\n")
        for i, line in enumerate(lines):
            if re.search(r"\bissue_(s|scores)\(", line) and "auto" not in line:          # call site of an issuing lambda
                seen += 1
                if "elect_one_sync()" not in line:
                    bad.append(f"{name}:{i + 1}")
            elif re.search(r"\bumma2?_(ss|ts)\(", line):
                for j in range(i - 1, -1, -1):                                          # the block (or lambda) this call sits in
                    if re.search(r"auto issue_\w+ = ", lines[j]):
                        break
                    if re.match(r"\s*if \(.*\) \{", lines[j]):
                        seen += 1
                        if "elect_one_sync()" not in lines[j]:
                            bad.append(f"{name}:{i + 1}")
                        break
    assert seen >= 8 and not bad, f"MMA issue guarded by a lane compare at {bad} ({seen} sites seen)"


def test_clock_sampler_reports_only_samples_of_the_timed_region():
    """bench.ClockSampler is started well before a 0.3 s timed region (nvidia-smi needs a few hundred ms to produce a sample): only the
    samples taken after begin() may reach the JSON line, and a region that caught none falls back to the last samples before it."""
    import sys
    import time
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    class _Proc:
        def terminate(self):
            pass
    s = bench.ClockSampler(0)
    s.proc = _Proc()
    now = time.time()
    row = lambda mhz, cap: ["0", str(mhz), "1965", "400.0", "Not Active", "Not Active", "Not Active", "Active" if cap else "Not Active"]
    s.rows = [(now - 2.0, row(300, False)), (now - 1.5, row(500, False))]
    s.t0 = now - 1.0
    s.rows += [(now - 0.5, row(1800, True)), (now - 0.2, row(1700, True))]
    out = s.stop()
    assert out["samples"] == 2 and out["sm_mhz"] == 1750.0 and out["reasons"] == ["sw_power_cap"]
    s2 = bench.ClockSampler(0)
    s2.proc = _Proc()
    s2.rows = [(now - 2.0, row(300, False)), (now - 1.5, row(500, False))]
    s2.t0 = now - 1.0
    out2 = s2.stop()
    assert out2["samples"] == 2 and out2["sm_mhz"] == 400.0          # nothing inside: the last samples before the region


def test_graphed_step_refuses_cpu_tensors():
    import torch
    from spokennlp_b200.graphs import GraphedStep
    from spokennlp_b200.lib import B200Error
    with pytest.raises(B200Error):
        GraphedStep(lambda x: x, (torch.zeros(2),))
