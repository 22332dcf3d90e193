"""Throughput of the other BASELINE.json configs on one B200 (they are parity cases, not the headline bench line):
  1  ditto: BERT-base forward, 16 x 128 tokens, output_hidden_states + output_attentions, layer-0/head-9 diagonal pooling
  3  alimeeting4mug: PoNet-base [2, 4096] forward + backward
  4  mmvts: projector + merge-attention / co-attention cross encoders, N = 300 clips, forward + backward
  5  sliding-window inference sweep: 2k..32k-token synthetic documents -> 512-token windows -> encoder + head argmax
Writes one JSON line per measurement to stdout (and gpurun_out/config_sweeps.jsonl)."""
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformers import BertConfig  # noqa: E402

from spokennlp_b200 import BertModel, ops  # noqa: E402
from spokennlp_b200.windows import build_windows, collate, synthetic_document  # noqa: E402

OUT = []


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def emit(**kw):
    OUT.append(kw)
    print(json.dumps(kw), flush=True)


base = dict(hidden_size=768, num_attention_heads=12, intermediate_size=3072, num_hidden_layers=12, vocab_size=30523,
            max_position_embeddings=512, type_vocab_size=2, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
torch.manual_seed(0)
bert = BertModel(BertConfig(**base)).cuda().eval()

# ---- config 1: ditto ------------------------------------------------------------------------------------------------
ids = torch.randint(1000, 30522, (16, 128)).cuda()
mask = torch.ones(16, 128, dtype=torch.long).cuda()


def ditto():
    with torch.no_grad():
        o = bert(ids, attention_mask=mask, output_hidden_states=True, output_attentions=True, return_dict=True)
        diag = torch.diagonal(o.attentions[0][:, 9], dim1=1, dim2=2)
        return (((o.hidden_states[0] + o.hidden_states[-1]) / 2.0) * mask[:, :, None] * diag[:, :, None]).sum(1)


t = timeit(ditto)
emit(config=1, workload="ditto BERT-base fwd 16x128 + hidden_states + attentions + diag pooling", seq_per_s=16 / t, ms=t * 1e3)

# ---- config 5: window sweep ------------------------------------------------------------------------------------------
W = torch.randn(2, 768, device="cuda") * 0.02
bb = torch.zeros(2, device="cuda")
eng = bert.b200_engine()
for T in (2048, 4096, 8192, 16384, 32768):
    sents, labs = synthetic_document(T, seed=T)
    wi, wm, wt, wl = collate(build_windows(sents, labs, 512), device="cuda")
    n = wi.shape[0]

    def infer():
        with torch.no_grad():
            kb, kl = ops.mask_to_bias(wm)
            x16, _, _, _, _ = eng.forward(wi.view(-1), None, None, None, kb, kl, n, 512, save=False)
            return ops.cls_head_fwd(x16, W, bb, want_argmax=True)[1]

    t = timeit(infer)
    emit(config=5, workload=f"sliding-window inference, {T}-token document", windows=n, seq_per_s=n / t, docs_per_s=1 / t, tokens_per_s=T / t,
         ms=t * 1e3)

# ---- config 3: PoNet --------------------------------------------------------------------------------------------------
from spokennlp_b200.windows import synthetic_segments as synth_segments  # noqa: E402
from spokennlp_b200.modeling_ponet import PoNetConfig, PoNetModel  # noqa: E402

pcfg = PoNetConfig(**{**{k: v for k, v in base.items()}, "max_position_embeddings": 4096})
ponet = PoNetModel(pcfg, add_pooling_layer=False).cuda().train()
pid = torch.randint(1000, 30522, (2, 4096)).cuda()
seg, pmask = synth_segments(2, 4096, seed=1, pad_from=[4096, 3900])
seg, pmask = seg.cuda(), pmask.cuda()


def ponet_step():
    for p in ponet.parameters():
        p.grad = None
    out = ponet(pid, attention_mask=pmask, segment_ids=seg, return_dict=True).last_hidden_state
    out.float().mean().backward()


t = timeit(ponet_step, iters=5, warm=2)
emit(config=3, workload="PoNet-base [2,4096] fwd+bwd (drop-in autograd path)", seq_per_s=2 / t, tokens_per_s=8192 / t, ms=t * 1e3)
with torch.no_grad():
    t = timeit(lambda: ponet(pid, attention_mask=pmask, segment_ids=seg, return_dict=True), iters=5, warm=2)
emit(config=3, workload="PoNet-base [2,4096] fwd", seq_per_s=2 / t, tokens_per_s=8192 / t, ms=t * 1e3)
from spokennlp_b200.graphs import GraphedStep  # noqa: E402


def ponet_gstep(i, a):
    for p in ponet.parameters():
        p.grad = None
    out = ponet(i, attention_mask=a, segment_ids=seg, return_dict=True).last_hidden_state
    loss = out.float().mean()
    loss.backward()
    return loss.detach()


pgs = GraphedStep(ponet_gstep, (pid, pmask))
t = timeit(lambda: pgs(pid, pmask), iters=10, warm=2)
emit(config=3, workload="PoNet-base [2,4096] fwd+bwd, CUDA graph", seq_per_s=2 / t, tokens_per_s=8192 / t, ms=t * 1e3)
del pgs

# ---- config 4: mmvts cross encoders -------------------------------------------------------------------------------
from spokennlp_b200.modeling_cross import CoAttentionEncoder, LinearProjector, MergeAttentionEncoder  # noqa: E402

conf = types.SimpleNamespace(hidden_size=768, num_cross_encoder_layers=1, num_cross_encoder_heads=12, intermediate_size=3072,
                             max_seq_length=2048, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, ce_kv_hidden_size=1536,
                             hidden_size_vis=3328, hidden_size_audio=768)
proj = LinearProjector(conf).cuda()
encs = (("ma", MergeAttentionEncoder(conf).cuda()), ("ca", CoAttentionEncoder(conf).cuda()))
for bs in (1, 16, 64):          # one video per step is ~400 launches over 300 token rows (launch-bound); batches show the kernels' rate
    tfeat, vfeat, afeat = torch.randn(bs, 300, 768).cuda(), torch.randn(bs, 300, 3328).cuda(), torch.randn(bs, 300, 768).cuda()
    cmask = torch.ones(bs, 300).cuda()
    cmask[:, 250:] = 0
    for name, enc in encs:
        def step():
            for m in (proj, enc):
                for p in m.parameters():
                    p.grad = None
            a, b, c = proj(tfeat, vfeat, afeat)
            t_, v_, a_ = enc(cmask, a, b, c)
            (t_.mean() + v_.mean() + a_.mean()).backward()
        t = timeit(step, iters=5, warm=2)
        emit(config=4, workload=f"mmvts projector + {name} cross encoder, N=300 clips, batch {bs}, fwd+bwd", samples_per_s=bs / t, ms=t * 1e3)
        if bs <= 16:                # the same step replayed from one CUDA graph (spokennlp_b200.graphs.GraphedStep)
            from spokennlp_b200.graphs import GraphedStep

            def gstep(tf, vf, af, cm, enc=enc):
                for m in (proj, enc):
                    for p in m.parameters():
                        p.grad = None
                a, b, c = proj(tf, vf, af)
                t_, v_, a_ = enc(cm, a, b, c)
                loss = t_.mean() + v_.mean() + a_.mean()
                loss.backward()
                return loss.detach()
            gs = GraphedStep(gstep, (tfeat, vfeat, afeat, cmask))
            t = timeit(lambda: gs(tfeat, vfeat, afeat, cmask), iters=10, warm=2)
            emit(config=4, workload=f"mmvts projector + {name} cross encoder, N=300 clips, batch {bs}, fwd+bwd, CUDA graph", samples_per_s=bs / t, ms=t * 1e3)

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/config_sweeps.jsonl", "w") as f:
    for r in OUT:
        f.write(json.dumps(r) + "\n")
