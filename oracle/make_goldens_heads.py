"""Mint golden vectors for the topic-segmentation loss heads from the REAL reference, run in the build container.

TEST INFRASTRUCTURE ONLY.  Run here (needs /root/reference and `transformers`):

    python oracle/make_goldens_heads.py        ->  tests/golden/ts_heads.pt

What is executed (no arithmetic of ours): `LossCalculator`, `CSSL`, `TSSP`, `EopPairCosineSimilarity`, `get_loss_fct` and
`BertWithDAForSentenceLabelingTopicSegmentation` imported from /root/reference/emnlp2023-topic_segmentation/src/models.
  * forward goldens: `LossCalculator.forward` / the wrapper's forward under no_grad (SURVEY §8c trap 1: their
    `loss = torch.tensor(0, requires_grad=True).to(device); loss += ...` cannot run in grad mode on the CPU);
  * gradient goldens: the same sub-modules the forward calls (classifier + get_loss_fct, cssl, tssp, eop_pair_cos_sim),
    called one by one in grad mode and summed with the config's weights exactly as loss_calculator.py:47-71 does.
Inputs, weights and the `random` seed used for CSSL's eop_list draws are stored next to the outputs.
"""
from __future__ import annotations

import os
import random
import sys

sys.dont_write_bytecode = True  # /root/reference is read-only

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
REF_SRC = "/root/reference/emnlp2023-topic_segmentation/src"

from oracle.bert_oracle import OracleConfig, random_state_dict  # noqa: E402  (weights only)
from oracle.ts_heads_oracle import synth_pair_batch  # noqa: E402  (inputs only)

KW = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, num_hidden_layers=2, vocab_size=128,
          max_position_embeddings=128, type_vocab_size=2)
CASES = {
    "full_matrix": dict(do_da_ts=True, do_tssp=True, ts_score_predictor="lt", focal_loss_gamma=0.0, weight_label_zero=0.5,
                        cl_loss_weight=0.5, cl_temp=0.1, cl_anchor_level="eop_matrix", tssp_loss_weight=1.0),
    "focal_list": dict(do_da_ts=True, do_tssp=True, ts_score_predictor="lt", focal_loss_gamma=2.0, weight_label_zero=0.3,
                       cl_loss_weight=0.5, cl_temp=0.1, cl_anchor_level="eop_list", cl_positive_k=1, cl_negative_k=3,
                       tssp_loss_weight=0.5),
    "cos_only": dict(do_da_ts=False, do_tssp=False, ts_score_predictor="cos", ts_score_predictor_cos_temp=0.5,
                     cl_loss_weight=0.0, tssp_loss_weight=0.0),
    # edge cases: an example without labels, an example with a single label, class weights without focal, ts weight != 1
    "ragged_weighted": dict(_ragged=True, do_da_ts=True, do_tssp=True, ts_score_predictor="lt", focal_loss_gamma=0.0,
                            weight_label_zero=0.7, ts_loss_weight=0.5, cl_loss_weight=0.25, cl_temp=0.2,
                            cl_anchor_level="eop_matrix", tssp_loss_weight=2.0),
}
DEFAULTS = dict(num_labels=2, classifier_dropout=None, do_da_ts=False, do_tssp=False, do_cssl=False, ts_score_predictor="lt",
                ts_score_predictor_cos_temp=1, focal_loss_gamma=0.0, weight_label_zero=0.5, ts_loss_weight=1.0,
                cl_loss_weight=0.0, tssp_loss_weight=0.0, cl_temp=1, cl_anchor_level="eop_matrix", cl_positive_k=1,
                cl_negative_k=1, num_tssp_labels=3)
RANDOM_SEED = 1234


def make_cfg(case):
    from transformers import BertConfig
    cfg = BertConfig(attn_implementation="eager", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **KW)
    for k, v in {**DEFAULTS, **case}.items():
        setattr(cfg, k, v)
    return cfg


def main():
    sys.path.insert(0, REF_SRC)
    from models.bert_for_ts import BertWithDAForSentenceLabelingTopicSegmentation as Wrapper
    from models.modules.loss_calculator import LossCalculator
    from models.modules.utils import get_loss_fct

    B, S, H = 3, 128, KW["hidden_size"]
    batch = synth_pair_batch(B, S, KW["vocab_size"], seed=17)
    sd = random_state_dict(OracleConfig(**KW), seed=7)
    g = torch.Generator().manual_seed(23)
    heads = dict(cls_w=torch.randn(2, H, generator=g) * 0.05, cls_b=torch.randn(2, generator=g) * 0.05,
                 tssp_w=torch.randn(3, H, generator=g) * 0.05, tssp_b=torch.randn(3, generator=g) * 0.05)
    h_rand = torch.randn(B, 2, S, H, generator=g)                  # stand-in encoder outputs for the heads-only goldens
    out = dict(config=KW, weight_seed=7, batch=batch, heads=heads, h_rand=h_rand, random_seed=RANDOM_SEED, cases={})

    # ragged variant of the same batch: example 1 carries no label at all in either view, example 2 exactly one
    ragged = {k: v.clone() for k, v in batch.items()}
    for v in range(2):
        ragged["labels"][1, v] = -100
        ragged["extract_eop_segment_ids"][1, v] = 0
        ragged["eop_index_for_aggregate_batch_eop_features"][1, v] = 0
        ragged["sent_token_mask"][1, v] = -100
        ragged["sent_pair_orders"][1, v] = -100
        first = int((ragged["labels"][2, v] != -100).nonzero()[0])
        keep = ragged["labels"][2, v, first].clone()
        ragged["labels"][2, v] = -100
        ragged["labels"][2, v, first] = keep
        ragged["extract_eop_segment_ids"][2, v] = 0
        ragged["extract_eop_segment_ids"][2, v, first] = 1
        ragged["eop_index_for_aggregate_batch_eop_features"][2, v] = 0
        ragged["eop_index_for_aggregate_batch_eop_features"][2, v, 1] = 1
    out["batch_ragged"] = ragged
    default_batch = batch

    for name, case in CASES.items():
        batch = ragged if case.get("_ragged") else default_batch
        case = {k: v for k, v in case.items() if not k.startswith("_")}
        cfg = make_cfg(case)
        lc = LossCalculator(cfg)
        with torch.no_grad():
            lc.classifier.weight.copy_(heads["cls_w"]); lc.classifier.bias.copy_(heads["cls_b"])
            lc.tssp.classifier.weight.copy_(heads["tssp_w"]); lc.tssp.classifier.bias.copy_(heads["tssp_b"])
        kw0 = dict(extract_eop_segment_ids=batch["extract_eop_segment_ids"][:, 0],
                   eop_index_for_aggregate_batch_eop_features=batch["eop_index_for_aggregate_batch_eop_features"][:, 0])
        kw1 = dict(sent_token_mask=batch["sent_token_mask"][:, 1], sent_pair_orders=batch["sent_pair_orders"][:, 1], da_example_flag=True,
                   extract_eop_segment_ids=batch["extract_eop_segment_ids"][:, 1],
                   eop_index_for_aggregate_batch_eop_features=batch["eop_index_for_aggregate_batch_eop_features"][:, 1])
        rec = {"case": case, "ragged": batch is ragged}
        # ---- heads only, forward (the reference's own LossCalculator.forward, no_grad)
        with torch.no_grad():
            random.seed(RANDOM_SEED)
            l0, lg0, cs0 = lc(sequence_output=h_rand[:, 0], labels=batch["labels"][:, 0], **kw0)
            l1, lg1, cs1 = lc(sequence_output=h_rand[:, 1], labels=batch["labels"][:, 1], **kw1)
        rec.update(anchor_loss=l0.clone(), anchor_logits=lg0.clone(), anchor_cos=cs0.clone(), da_loss=l1.clone(), da_logits=lg1.clone())
        # ---- heads only, gradients (sub-modules in grad mode, weighted as loss_calculator.py:47-71)
        for p in lc.parameters():
            p.grad = None
        h0 = h_rand[:, 0].clone().requires_grad_(True)
        h1 = h_rand[:, 1].clone().requires_grad_(True)
        random.seed(RANDOM_SEED)

        def ts(h, labels):
            if cfg.ts_score_predictor == "lt":
                logits = lc.classifier(h)
                return get_loss_fct(gamma=cfg.focal_loss_gamma, weight_label_zero=cfg.weight_label_zero, device=h.device)(
                    logits.reshape(-1, cfg.num_labels), labels.reshape(-1))
            cos, lab = lc.eop_pair_cos_sim(h, labels)
            return torch.nn.BCEWithLogitsLoss()(cos.reshape(-1), lab.reshape(-1).float())
        tot0 = cfg.ts_loss_weight * ts(h0, batch["labels"][:, 0])
        if cfg.cl_loss_weight != 0:
            tot0 = tot0 + cfg.cl_loss_weight * lc.cssl(sequence_output=h0, labels=batch["labels"][:, 0], **kw0)
        tot1 = cfg.ts_loss_weight * ts(h1, batch["labels"][:, 1])
        if cfg.tssp_loss_weight != 0:
            tot1 = tot1 + cfg.tssp_loss_weight * lc.tssp(sent_token_mask=kw1["sent_token_mask"], da_seq_output=h1,
                                                         da_sent_pair_orders=kw1["sent_pair_orders"])
        assert abs(float(tot0) - float(l0)) < 1e-5 * max(1.0, abs(float(l0))), (float(tot0), float(l0))
        assert abs(float(tot1) - float(l1)) < 1e-5 * max(1.0, abs(float(l1))), (float(tot1), float(l1))
        (tot0 + tot1).backward()
        rec.update(grad_h0=h0.grad.clone(), grad_h1=h1.grad.clone(),
                   grad_cls_w=None if lc.classifier.weight.grad is None else lc.classifier.weight.grad.clone(),
                   grad_cls_b=None if lc.classifier.bias.grad is None else lc.classifier.bias.grad.clone(),
                   grad_tssp_w=None if lc.tssp.classifier.weight.grad is None else lc.tssp.classifier.weight.grad.clone(),
                   grad_tssp_b=None if lc.tssp.classifier.bias.grad is None else lc.tssp.classifier.bias.grad.clone())
        # ---- the wrapper end to end (HF BertModel inside, forward only)
        wr = Wrapper(cfg).eval()
        wr.bert.load_state_dict(sd, strict=False)
        with torch.no_grad():
            wr.loss_calculator.classifier.weight.copy_(heads["cls_w"]); wr.loss_calculator.classifier.bias.copy_(heads["cls_b"])
            wr.loss_calculator.tssp.classifier.weight.copy_(heads["tssp_w"]); wr.loss_calculator.tssp.classifier.bias.copy_(heads["tssp_b"])
            random.seed(RANDOM_SEED)
            wl, wlogits, wcos = wr(batch["input_ids"], attention_mask=batch["attention_mask"], token_type_ids=batch["token_type_ids"],
                                   labels=batch["labels"], extract_eop_segment_ids=batch["extract_eop_segment_ids"],
                                   eop_index_for_aggregate_batch_eop_features=batch["eop_index_for_aggregate_batch_eop_features"],
                                   sent_token_mask=batch["sent_token_mask"], sent_pair_orders=batch["sent_pair_orders"])[:3]
        rec.update(wrapper_loss=wl.clone(), wrapper_logits=wlogits.clone(), wrapper_cos=wcos.clone())
        out["cases"][name] = rec
        print(f"{name}: heads loss {float(l0):.6f} + {float(l1):.6f}, wrapper loss {float(wl):.6f}")
    out["source"] = "transformers %s + /root/reference emnlp2023-topic_segmentation/src/models" % __import__("transformers").__version__
    torch.save(out, os.path.join(OUT, "ts_heads.pt"))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    main()
