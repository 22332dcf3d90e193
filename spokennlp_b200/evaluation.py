"""Prediction writer and example-level metric of the topic-segmentation path.

What it mirrors (emnlp2023-topic_segmentation/src/):
  * ts_sentence_seq_labeling.py:1138-1177   per-window logits -> argmax at the labelled [BOS] rows (or the cosine predictor's
                                            thresholded similarities) -> string labels, logits kept beside them
  * ts_sentence_seq_labeling.py:1184-1201   windows regrouped by example id (a document is several windows), one JSON line per
                                            example with sentences / labels / int_labels / predictions / predict_logits
                                            (/ eop_pair_cos_sim)
  * ts_sentence_seq_labeling.py:1203-1216   examples with no labelled sentence are dropped before the example-level metric
  * metrics/seqeval.py:172-238              `compute_window_metric`: 0/1 boundary strings -> segment masses -> Pk and
                                            WindowDiff per example (averaged), binary precision / recall / F1 over all sentences
  * metrics/seqeval.py:248-373              `compute_metric_example_level`: argmax metric (seqeval's entity-level P/R/F1 on the
                                            'B-EOP' / 'O' tags), threshold / top-k / top-k-with-threshold / F1@k variants

Third-party arithmetic the reference delegates to and this file restates (neither package exists in this image):
  * `seqeval.metrics.classification_report` on the tag set ['B-EOP', 'O']: every 'B-EOP' tag is a one-token entity of type
    EOP and 'O' is outside, so entity-level precision / recall / F1 equal the token-level scores of class 'B-EOP'; the
    reference's `overall_accuracy` is plain token accuracy.
  * `segeval.window.pk.pk` / `segeval.window.windowdiff.window_diff` (segeval 2.0.11, mass format, default arguments):
    window size k = round(mean reference segment mass / 2) with Python's round-half-even on a Decimal, at least 2;
    Pk = fraction of the N-k probes (i, i+k) on which "same segment?" differs between reference and hypothesis
    (Beeferman et al. 1999); WindowDiff = fraction of the N-k windows of k+1 units whose boundary COUNT differs (Pevzner &
    Hearst 2002).  Restated from the published definitions: segeval itself is absent, so Pk / WD are unpinned against it
    (tests/test_evaluation.py checks them against brute-force statements of the two definitions).

The inputs are what the device path produces: `logits` [n_windows, S, C] (or the argmax the cls-head kernel emits) and the
label tensor; everything here is host-side bookkeeping over O(labelled sentences) values.
"""
from __future__ import annotations

import json
from decimal import Decimal
from types import SimpleNamespace
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

LABEL_LIST = ["B-EOP", "O"]          # ts_sentence_seq_labeling.py:180 — class 0 = topic boundary AFTER this sentence


# ------------------------------------------------------------------------------------------------ segment arithmetic
def masses_from_boundaries(labels: Sequence[int]) -> List[int]:
    """[1, 1, 0, 0, 1, 1] -> [1, 1, 3, 1]: 1 = this sentence ENDS a segment (seqeval.py:177-190)."""
    mass, cur = [], 0
    for v in labels:
        cur += 1
        if v == 1:
            mass.append(cur)
            cur = 0
    if cur > 0:
        mass.append(cur)
    return mass


def positions_from_masses(masses: Sequence[int]) -> List[int]:
    """(2, 3) -> [0, 0, 1, 1, 1]: segment index of every unit."""
    out: List[int] = []
    for i, m in enumerate(masses):
        out.extend([i] * int(m))
    return out


def window_size(ref_masses: Sequence[int]) -> int:
    """segeval's default: half the mean reference segment mass, rounded half-to-even, never below 2."""
    k = int(round(Decimal(sum(ref_masses)) / len(ref_masses) / 2))
    return k if k > 1 else 2


def pk(hyp_masses: Sequence[int], ref_masses: Sequence[int], k: Optional[int] = None) -> float:
    ref, hyp = positions_from_masses(ref_masses), positions_from_masses(hyp_masses)
    if len(ref) != len(hyp):
        raise ValueError("segmentations cover different numbers of units")
    k = window_size(ref_masses) if k is None else k
    n = len(ref) - k
    if n <= 0:
        return 0.0
    diff = sum((ref[i] == ref[i + k]) != (hyp[i] == hyp[i + k]) for i in range(n))
    return diff / n


def window_diff(hyp_masses: Sequence[int], ref_masses: Sequence[int], k: Optional[int] = None) -> float:
    ref, hyp = positions_from_masses(ref_masses), positions_from_masses(hyp_masses)
    if len(ref) != len(hyp):
        raise ValueError("segmentations cover different numbers of units")
    k = window_size(ref_masses) if k is None else k
    n = len(ref) - k
    if n <= 0:
        return 0.0
    # boundaries inside the window of units i .. i+k = segment-index difference of its ends (indices are non-decreasing)
    diff = sum((ref[i + k] - ref[i]) != (hyp[i + k] - hyp[i]) for i in range(n))
    return diff / n


def _prf(tp: int, n_pred: int, n_true: int):
    p = tp / n_pred if n_pred else 0.0
    r = tp / n_true if n_true else 0.0
    f = 2 * p * r / (p + r) if (p + r) else 0.0
    return p, r, f


def compute_window_metric(predictions: Sequence[Sequence[int]], references: Sequence[Sequence[int]], prefix: str = "") -> Dict[str, float]:
    """seqeval.py:172-238.  predictions / references: per example, 1 = the sentence ends a topic."""
    one_minus_pk, one_minus_wd = [], []
    for y_pred, y_true in zip(predictions, references):
        pm, tm = masses_from_boundaries(y_pred), masses_from_boundaries(y_true)
        if not tm or sum(pm) != sum(tm):      # the reference swallows such examples (bare `except`, seqeval.py:214-215)
            continue
        one_minus_pk.append(1 - pk(pm, tm))
        one_minus_wd.append(1 - window_diff(pm, tm))
    r_pk = round(float(np.mean(one_minus_pk)), 4) if one_minus_pk else float("nan")
    r_wd = round(float(np.mean(one_minus_wd)), 4) if one_minus_wd else float("nan")
    flat_p = [int(v) for ex in predictions for v in ex]
    flat_t = [int(v) for ex in references for v in ex]
    tp = sum(1 for a, b in zip(flat_p, flat_t) if a == 1 and b == 1)
    p, r, f1 = _prf(tp, sum(flat_p), sum(flat_t))
    # (the reference also derives micro-F1 and the mean boundary counts per example here but does not return them)
    return {
        prefix + "1-pk": r_pk,
        prefix + "1-wd": r_wd,
        prefix + "precision": round(p, 4),
        prefix + "recall": round(r, 4),
        prefix + "f1": round(f1, 4),
        prefix + "pk": 1 - r_pk,
        prefix + "wd": 1 - r_wd,
    }


def tag_metric(predictions: Sequence[Sequence[str]], references: Sequence[Sequence[str]]) -> Dict[str, float]:
    """`metric.compute` of the reference on ['B-EOP', 'O'] tags (seqeval.py:126-170): entity-level scores of type EOP, which for
    one-token entities are the token-level scores of the 'B-EOP' class, plus token accuracy."""
    tp = n_pred = n_true = right = total = 0
    for pr, rf in zip(predictions, references):
        for a, b in zip(pr, rf):
            total += 1
            right += a == b
            n_pred += a == "B-EOP"
            n_true += b == "B-EOP"
            tp += a == "B-EOP" and b == "B-EOP"
    p, r, f1 = _prf(tp, n_pred, n_true)
    return {"EOP": {"precision": p, "recall": r, "f1": f1, "number": n_true}, "overall_precision": p, "overall_recall": r,
            "overall_f1": f1, "overall_accuracy": right / max(total, 1)}


def _softmax_rows(x: np.ndarray) -> np.ndarray:
    x = x - x.max(axis=-1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=-1, keepdims=True)


def compute_metric_example_level(predictions_logits, labels, label_list=LABEL_LIST, custom_args=None, data_args=None,
                                 reverse_logits: bool = False, ts_score_predictor: str = "lt", mode: str = "test") -> Dict[str, float]:
    """seqeval.py:248-373.  `predictions_logits`: per example, [n_sentences, C] logits ("lt") or [n_sentences] sigmoid(cos)
    scores ("cos"); `labels`: per example, int labels (0 = boundary).  `custom_args`: threshold / topk / topk_with_threshold /
    f1_at_k as in arguments.py; `data_args.return_entity_level_metrics`."""
    custom_args = custom_args or SimpleNamespace(threshold=None, topk=None, topk_with_threshold=False, f1_at_k=None)
    ret_entity = bool(getattr(data_args, "return_entity_level_metrics", False))
    if ts_score_predictor == "lt":
        predictions = [np.argmax(np.asarray(lg, dtype=np.float64), axis=-1) for lg in predictions_logits]
        scores = [_softmax_rows(np.asarray(lg, dtype=np.float64))[:, 0] for lg in predictions_logits]
    else:
        predictions = [(np.asarray(lg) > 0.5).astype(np.int32).tolist() for lg in predictions_logits]
        scores = [[1 - v for v in lg] for lg in predictions_logits]
    if reverse_logits:
        predictions = [[1 - v for v in pred] for pred in predictions]
    keep = lambda seq, lab: [v for v, l in zip(seq, lab) if l != -100]
    true_predictions = [[label_list[int(p)] for p in keep(pr, lb)] for pr, lb in zip(predictions, labels)]
    true_labels = [[label_list[int(l)] for l in keep(lb, lb)] for lb in labels]
    results = tag_metric(true_predictions, true_labels)
    results["accuracy"] = results["overall_accuracy"]
    true_bin = [[int(not l) for l in keep(lb, lb)] for lb in labels]              # 1 = the sentence ends a topic
    custom: Dict[str, float] = {}
    thr = getattr(custom_args, "threshold", None)
    if thr is not None:
        ge = (lambda v: v >= thr) if ts_score_predictor == "lt" else (lambda v: v > thr)
        pred_bin = [[1 if ge(v) else 0 for v in keep(sc, lb)] for sc, lb in zip(scores, labels)]
        custom.update(compute_window_metric(pred_bin, true_bin, prefix=f"threshold_{thr}_example_level_"))
    topk = getattr(custom_args, "topk", None)
    if topk is not None:
        prefix = f"topk_{topk}_example_level_"
        para = [keep(sc, lb) for sc, lb in zip(scores, labels)]
        ranked = [sorted([(v, i) for i, v in enumerate(ex)], reverse=True) for ex in para]

        def pick(cond):
            out = []
            for ex, rk in zip(para, ranked):
                row = [0] * len(ex)
                for v, i in rk[:topk]:
                    if cond(v):
                        row[i] = 1
                out.append(row)
            return out
        res = compute_window_metric(pick(lambda v: True), true_bin, prefix=prefix)
        # the reference indexes the (k+1)-th score and raises when an example has no more than k sentences (seqeval.py:318);
        # examples that short are skipped here instead
        kth = [rk[topk][0] for rk in ranked if len(rk) > topk]
        res[prefix + "kth_scores_avg"] = round(float(sum(kth) / len(kth)), 3) if kth else float("nan")
        custom.update(res)
        if getattr(custom_args, "topk_with_threshold", False):
            assert thr is not None
            custom.update(compute_window_metric(pick(lambda v: v >= thr), true_bin, prefix=f"topk_{topk}_with_threshold_{thr}_example_level_"))
    k_at = getattr(custom_args, "f1_at_k", None)
    if k_at:
        pred_bin = [[1 if v >= thr else 0 for v in keep(sc, lb)] for sc, lb in zip(scores, labels)]
        soft = []
        for pred, lab in zip(pred_bin, true_bin):
            for i, p in enumerate(pred):
                if p == 0 or lab[i] == 1:
                    continue
                for j in range(max(0, i - k_at), min(len(pred) - 1, i + k_at) + 1):   # a near miss counts as the nearby true boundary
                    if lab[j] == 1:
                        pred[i], pred[j] = 0, 1
                        break
            soft.append(pred)
        custom.update(compute_window_metric(soft, true_bin, prefix=f"f1@{k_at}_example_level_"))
    final: Dict[str, float] = {}
    if ret_entity:
        for key, value in results.items():
            if isinstance(value, dict):
                for n, v in value.items():
                    final[f"{key}_{n}"] = v
            else:
                final[key] = value
    else:
        final.update({"precision": results["overall_precision"], "recall": results["overall_recall"], "f1": results["overall_f1"]})
    final.update(custom)
    return final


# ------------------------------------------------------------------------------------------------ prediction writer
def window_predictions(logits, labels, label_list=LABEL_LIST, ts_score_predictor: str = "lt", argmax=None):
    """ts_sentence_seq_labeling.py:1138-1164 for the anchor view.  `logits`: [n_windows, S, C] ("lt") or [n_windows, max_eop]
    sigmoid(cos) scores ("cos"); `labels`: [n_windows, S] with -100 off the labelled [BOS] rows; `argmax` (optional, "lt"):
    the int32 [n_windows, S] the cls-head kernel already produced.  Returns per window: string predictions, string labels,
    int labels, and the logits of the labelled rows."""
    logits, labels = np.asarray(logits), np.asarray(labels)
    true_labels = [[label_list[int(l)] for l in row if l != -100] for row in labels]
    true_int = [[int(l) for l in row if l != -100] for row in labels]
    if ts_score_predictor == "lt":
        am = np.asarray(argmax) if argmax is not None else np.argmax(logits, axis=2)
        preds = [[label_list[int(p)] for p, l in zip(pr, lb) if l != -100] for pr, lb in zip(am, labels)]
        plog = [[p.tolist() for p, l in zip(lg, lb) if l != -100] for lg, lb in zip(logits, labels)]
    elif ts_score_predictor == "cos":
        pb = (logits > 0.5).astype(np.int32)
        preds = [[label_list[int(p)] for p in row[:len(ti)]] for row, ti in zip(pb, true_int)]
        plog = [row[:len(ti)].tolist() for row, ti in zip(logits, true_int)]
    else:
        raise ValueError(f"not supported ts_score_predictor {ts_score_predictor}")
    return preds, true_labels, true_int, plog


def merge_windows(example_ids: Sequence[int], num_examples: int, sentences: Sequence[Sequence[str]], preds, true_labels, true_int, plog,
                  eop_pair_cos_sim=None) -> List[dict]:
    """ts_sentence_seq_labeling.py:1184-1199: windows appended to their example in dataset order (the sentence two
    consecutive windows share is listed twice in `sentences`, as upstream; its label is scored once — Appendix A.2)."""
    out = [{"sentences": [], "labels": [], "int_labels": [], "predictions": [], "predict_logits": []} for _ in range(num_examples)]
    for pr, sl, lb, il, eid, pl in zip(preds, sentences, true_labels, true_int, example_ids, plog):
        o = out[eid]
        o["sentences"].extend(sl)
        o["labels"].extend(lb)
        o["predictions"].extend(pr)
        o["predict_logits"].extend(pl)
        o["int_labels"].extend(int(v) for v in il)
    if eop_pair_cos_sim is not None:
        for o in out:
            o["eop_pair_cos_sim"] = []
        for eid, cs in zip(example_ids, eop_pair_cos_sim):
            out[eid]["eop_pair_cos_sim"].extend(float(v) for v in cs if v != -100)
    return out


def write_predictions(path: str, records: Iterable[dict]) -> None:
    with open(path, "w") as fh:
        fh.writelines(json.dumps(r, ensure_ascii=False) + "\n" for r in records)


def example_level_metric(records: Sequence[dict], label_list=LABEL_LIST, custom_args=None, data_args=None,
                         ts_score_predictor: str = "lt", mode: str = "predict") -> Dict[str, float]:
    """ts_sentence_seq_labeling.py:1203-1218: drop examples without labelled sentences, then the example-level metric."""
    lg = [r["predict_logits"] for r in records if len(r["int_labels"])]
    lb = [r["int_labels"] for r in records if len(r["int_labels"])]
    res = compute_metric_example_level(lg, lb, label_list, custom_args, data_args, ts_score_predictor=ts_score_predictor, mode=mode)
    res[f"{mode}_examples"] = len(records)
    return res
