"""Sliding-window packer (spokennlp_b200/windows.py) against the properties the reference's rule guarantees
(ts_sentence_seq_labeling.py:811-873)."""
import torch

from spokennlp_b200.windows import IGNORE, build_windows, collate, synthetic_document

BOS, CLS, S = 30522, 101, 512


def _check(T, seed):
    sents, labs = synthetic_document(T, seed)
    wins = build_windows(sents, labs, S, cls_id=CLS)
    assert wins
    scored = {}
    for wi, w in enumerate(wins):
        assert len(w.input_ids) == len(w.attention_mask) == len(w.labels) == S
        assert w.input_ids[0] == CLS and w.labels[0] == IGNORE
        n = sum(w.attention_mask)
        assert all(m == 1 for m in w.attention_mask[:n]) and all(m == 0 for m in w.attention_mask[n:])
        # labels sit exactly on [BOS] positions; the window's last sentence is masked
        bos_pos = [p for p in range(n) if w.input_ids[p] == BOS]
        lab_pos = [p for p in range(S) if w.labels[p] != IGNORE]
        assert set(lab_pos) <= set(bos_pos)
        assert len(lab_pos) == max(0, len(bos_pos) - 1) and (not bos_pos or w.labels[bos_pos[-1]] == IGNORE)
        sent_ids = list(w.sent_range)
        for p, sid in zip(bos_pos[:-1], sent_ids):
            assert w.labels[p] == labs[sid]
            scored[sid] = scored.get(sid, 0) + 1
        if wi > 0:   # consecutive windows share exactly one sentence unless the previous one held a single sentence
            prev = wins[wi - 1].sent_range
            assert w.sent_range.start in (prev.stop - 1, prev.stop)
    # every sentence except the last of the document is scored exactly once
    assert all(scored.get(i, 0) == 1 for i in range(len(sents) - 1)), [i for i in range(len(sents) - 1) if scored.get(i, 0) != 1][:5]
    assert (len(sents) - 1) not in scored
    return wins


def test_window_rule_properties_over_lengths():
    for T, seed in ((300, 0), (2048, 1), (8192, 2), (32768, 3)):
        wins = _check(T, seed)
        # ~ceil(T / (S - mean sentence length)) rows, almost all full
        assert abs(len(wins) - T / (S - 26)) < 0.15 * len(wins) + 2
        full = sum(1 for w in wins if sum(w.attention_mask) >= S - 45)
        assert full >= len(wins) - 1


def test_single_overlong_sentence_is_truncated_and_masked():
    sents = [[BOS] + [2000] * 700, [BOS] + [2001] * 10]
    wins = build_windows(sents, [1, 0], S, cls_id=CLS)
    assert len(wins) == 2 and sum(wins[0].attention_mask) == S
    assert all(l == IGNORE for l in wins[0].labels) and all(l == IGNORE for l in wins[1].labels)


def test_collate_shapes():
    sents, labs = synthetic_document(3000, 5)
    ids, mask, tt, labels = collate(build_windows(sents, labs, S))
    assert ids.shape == mask.shape == tt.shape == labels.shape and ids.shape[1] == S and ids.dtype == torch.long


def test_synthetic_segments_follow_the_driver_convention():
    from spokennlp_b200.windows import synthetic_segments
    seg, mask = synthetic_segments(2, 256, seed=3, pad_from=[256, 200])
    assert seg[:, 0].tolist() == [0, 0] and mask[0].all() and mask[1, :200].all() and not mask[1, 200:].any()
    for b, end in ((0, 256), (1, 200)):
        body = seg[b, 1:end]
        assert bool((body[1:] - body[:-1] >= 0).all()) and int(body[0]) == 1          # monotone runs numbered from 1
        runs = torch.unique_consecutive(body, return_counts=True)[1]
        assert int(runs[:-1].min()) >= 8 and int(runs.max()) <= 40
        assert bool((seg[b, end:] == int(body.max()) + 1).all())
