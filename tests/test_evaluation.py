"""Prediction writer + example-level metric (spokennlp_b200/evaluation.py; reference: ts_sentence_seq_labeling.py:1138-1222,
metrics/seqeval.py:172-373).  P / R / F1 are pinned against scikit-learn (the library the reference calls, present here); Pk and
WindowDiff against brute-force statements of the published definitions (segeval is absent: unpinned against the package)."""
import json
import random
from types import SimpleNamespace

import numpy as np
import pytest

from spokennlp_b200 import evaluation as E


def test_masses_follow_the_reference_docstring_example():
    assert E.masses_from_boundaries([1, 1, 0, 0, 1, 1]) == [1, 1, 3, 1]            # seqeval.py:178-179
    assert E.masses_from_boundaries([0, 0, 0]) == [3] and E.masses_from_boundaries([]) == []
    assert E.positions_from_masses([2, 3]) == [0, 0, 1, 1, 1]


def _brute_pk(hyp, ref, k):
    n = len(ref)
    diff = 0
    for i in range(n - k):
        same_ref = all(ref[j] == 0 for j in range(i, i + k))      # no reference boundary between unit i and unit i+k
        same_hyp = all(hyp[j] == 0 for j in range(i, i + k))
        diff += same_ref != same_hyp
    return diff / (n - k)


def _brute_wd(hyp, ref, k):
    n = len(ref)
    return sum(sum(ref[i:i + k]) != sum(hyp[i:i + k]) for i in range(n - k)) / (n - k)


def test_pk_and_windowdiff_match_the_definitions():
    rng = random.Random(0)
    for _ in range(200):
        n = rng.randint(6, 60)
        ref = [int(rng.random() < 0.25) for _ in range(n)]
        hyp = [int(rng.random() < 0.3) for _ in range(n)]
        ref[-1] = hyp[-1] = 1                                   # the last sentence always ends a segment
        rm, hm = E.masses_from_boundaries(ref), E.masses_from_boundaries(hyp)
        k = E.window_size(rm)
        assert k >= 2 and sum(rm) == sum(hm) == n
        if n - k <= 0:
            continue
        assert abs(E.pk(hm, rm) - _brute_pk(hyp, ref, k)) < 1e-12
        assert abs(E.window_diff(hm, rm) - _brute_wd(hyp, ref, k)) < 1e-12
        assert E.pk(rm, rm) == 0.0 and E.window_diff(rm, rm) == 0.0
        assert E.window_diff(hm, rm) >= E.pk(hm, rm) - 1e-12     # WindowDiff penalises everything Pk does (Pevzner & Hearst)


def test_window_size_rounds_half_to_even_like_decimal_round():
    assert E.window_size([5, 5]) == 2          # 2.5 -> 2
    assert E.window_size([7, 7]) == 4          # 3.5 -> 4
    assert E.window_size([1, 1, 1]) == 2       # never below 2
    assert E.window_size([10, 12]) == 6        # 5.5 -> 6


def test_binary_prf_match_sklearn():
    from sklearn.metrics import f1_score, precision_score, recall_score
    rng = np.random.default_rng(1)
    preds = [rng.integers(0, 2, n).tolist() for n in (12, 30, 7, 19)]
    refs = [rng.integers(0, 2, n).tolist() for n in (12, 30, 7, 19)]
    for p, r in zip(preds, refs):
        p[-1] = r[-1] = 1
    res = E.compute_window_metric(preds, refs, prefix="x_")
    fp, fr = sum(preds, []), sum(refs, [])
    assert res["x_precision"] == round(precision_score(fr, fp), 4)
    assert res["x_recall"] == round(recall_score(fr, fp), 4)
    assert res["x_f1"] == round(f1_score(fr, fp), 4)
    assert set(res) == {"x_1-pk", "x_1-wd", "x_precision", "x_recall", "x_f1", "x_pk", "x_wd"}      # seqeval.py:229-237
    assert abs(res["x_pk"] - (1 - res["x_1-pk"])) < 1e-12


def test_tag_metric_is_the_b_eop_class_score():
    from sklearn.metrics import precision_recall_fscore_support
    rng = np.random.default_rng(2)
    pr = [[E.LABEL_LIST[i] for i in rng.integers(0, 2, 25)] for _ in range(4)]
    rf = [[E.LABEL_LIST[i] for i in rng.integers(0, 2, 25)] for _ in range(4)]
    res = E.tag_metric(pr, rf)
    p, r, f, _ = precision_recall_fscore_support([t == "B-EOP" for t in sum(rf, [])], [t == "B-EOP" for t in sum(pr, [])], average="binary")
    assert abs(res["overall_precision"] - p) < 1e-12 and abs(res["overall_recall"] - r) < 1e-12 and abs(res["overall_f1"] - f) < 1e-12
    assert res["EOP"]["number"] == sum(t == "B-EOP" for t in sum(rf, []))


def _windows(seed=0, n_windows=7, S=48, n_examples=3):
    rng = np.random.default_rng(seed)
    logits = rng.normal(size=(n_windows, S, 2)).astype(np.float32)
    labels = np.full((n_windows, S), -100, dtype=np.int64)
    sentences = []
    for w in range(n_windows):
        pos = np.arange(1, S - 2, 6)[: rng.integers(2, 7)]
        labels[w, pos] = rng.integers(0, 2, len(pos))
        sentences.append([f"w{w}s{i}" for i in range(len(pos) + 1)])
    example_ids = sorted(rng.integers(0, n_examples, n_windows).tolist())
    return logits, labels, sentences, example_ids


def test_prediction_writer_regroups_windows_by_example(tmp_path):
    logits, labels, sentences, eids = _windows()
    preds, tl, ti, plog = E.window_predictions(logits, labels)
    # the kernel-side argmax may be handed in instead of being recomputed
    preds2, *_ = E.window_predictions(logits, labels, argmax=np.argmax(logits, axis=2).astype(np.int32))
    assert preds == preds2
    for w in range(len(labels)):
        keep = labels[w] != -100
        assert ti[w] == labels[w][keep].tolist()
        assert preds[w] == [E.LABEL_LIST[i] for i in logits[w][keep].argmax(-1)]
        assert np.allclose(plog[w], logits[w][keep])
    cos = np.full((len(labels), 10), -100.0)
    for w in range(len(labels)):
        cos[w, :len(ti[w])] = 0.25
    recs = E.merge_windows(eids, 3, sentences, preds, tl, ti, plog, eop_pair_cos_sim=cos)
    assert sum(len(r["int_labels"]) for r in recs) == int((labels != -100).sum())
    for e in range(3):
        want = [l for w, eid in enumerate(eids) if eid == e for l in ti[w]]
        assert recs[e]["int_labels"] == want and len(recs[e]["predictions"]) == len(want) == len(recs[e]["eop_pair_cos_sim"])
    path = tmp_path / "predict.txt"
    E.write_predictions(str(path), recs)
    back = [json.loads(l) for l in open(path)]
    assert back == json.loads(json.dumps(recs)) and set(back[0]) == {"sentences", "labels", "int_labels", "predictions", "predict_logits", "eop_pair_cos_sim"}


def test_example_level_metric_variants():
    logits, labels, sentences, eids = _windows(seed=3, n_windows=12, S=96, n_examples=4)
    preds, tl, ti, plog = E.window_predictions(logits, labels)
    recs = E.merge_windows(eids, 5, sentences, preds, tl, ti, plog)           # example 4 has no window: dropped by the metric
    args = SimpleNamespace(threshold=0.5, topk=3, topk_with_threshold=True, f1_at_k=1)
    res = E.example_level_metric(recs, custom_args=args, data_args=SimpleNamespace(return_entity_level_metrics=False))
    assert res["predict_examples"] == 5
    for key in ("precision", "recall", "f1", "threshold_0.5_example_level_pk", "topk_3_example_level_wd", "topk_3_example_level_kth_scores_avg",
                "topk_3_with_threshold_0.5_example_level_f1", "f1@1_example_level_f1"):
        assert key in res, key
    # softmax(logits)[:, 0] >= 0.5 is the argmax rule (ties aside), so the thresholded boundary F1 equals the tag F1
    assert abs(res["threshold_0.5_example_level_f1"] - round(res["f1"], 4)) < 1e-4
    assert res["f1@1_example_level_f1"] >= res["threshold_0.5_example_level_f1"]           # near misses are forgiven
    ent = E.example_level_metric(recs, custom_args=args, data_args=SimpleNamespace(return_entity_level_metrics=True))
    assert "EOP_f1" in ent and "overall_accuracy" in ent
    # perfect predictions
    perfect = [dict(r, predict_logits=[[5.0, -5.0] if l == 0 else [-5.0, 5.0] for l in r["int_labels"]]) for r in recs]
    pres = E.example_level_metric(perfect, custom_args=SimpleNamespace(threshold=0.5, topk=None, topk_with_threshold=False, f1_at_k=None))
    assert pres["f1"] == 1.0 and pres["threshold_0.5_example_level_pk"] == 0 and pres["threshold_0.5_example_level_1-wd"] == 1.0


def test_cos_predictor_path():
    scores = np.array([[0.9, 0.2, 0.7, -100.0], [0.1, 0.8, -100.0, -100.0]])
    labels = np.full((2, 16), -100)
    labels[0, [1, 5, 9]] = [1, 0, 1]
    labels[1, [1, 5]] = [0, 1]
    preds, tl, ti, plog = E.window_predictions(scores, labels, ts_score_predictor="cos")
    assert preds == [["O", "B-EOP", "O"], ["B-EOP", "O"]] and plog == [[0.9, 0.2, 0.7], [0.1, 0.8]]
    res = E.compute_metric_example_level(plog, ti, ts_score_predictor="cos", custom_args=SimpleNamespace(threshold=0.5, topk=None, topk_with_threshold=False, f1_at_k=None))
    assert res["f1"] == 1.0
    with pytest.raises(ValueError):
        E.window_predictions(scores, labels, ts_score_predictor="nope")
