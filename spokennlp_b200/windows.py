"""Sliding-window packing of long documents into fixed-length encoder rows (SURVEY.md §5 "long-context", Appendix A.2).

Restates the reference's window rule so that synthetic long documents (BASELINE config 5: 2k-32k tokens) can be fed to
the encoder exactly as the reference would feed them:
  emnlp2023-topic_segmentation/src/ts_sentence_seq_labeling.py:811-859 (identical logic at
  alimeeting4mug/src/topic_segment/ponet_topic_segmentation.py:615-661 and mmvts/src/main_multimodal.py:417-507).

A document is a list of sentences (token-id lists, each already prefixed by its [BOS] token) with one label per sentence.
Windows are cut at sentence ends, hold at most `max_seq_length` tokens including [CLS], consecutive windows overlap by
exactly one sentence, and the label of the last sentence of every window is masked (-100) because its successor is not
visible.  Windows are independent rows: the encoder sees them as a plain batch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import torch

IGNORE = -100


@dataclass
class Window:
    input_ids: List[int]
    attention_mask: List[int]
    labels: List[int]
    sent_range: range            # sentences (document indices) present in this window


def build_windows(sentences: Sequence[Sequence[int]], sent_labels: Sequence[int], max_seq_length: int, cls_id: int = 101,
                  pad_id: int = 0, bos_position_label: bool = True) -> List[Window]:
    """ts_sentence_seq_labeling.py:811-873.  `sentences[i][0]` is the [BOS] token that carries sentence i's label."""
    tokens: List[int] = []
    labels: List[int] = []
    ends: List[int] = []                       # index of the last token of sentence i in `tokens` (:812)
    for sent, lab in zip(sentences, sent_labels):
        tokens.extend(sent)
        labels.extend([lab if bos_position_label else IGNORE] + [IGNORE] * (len(sent) - 1))
        ends.append(len(tokens) - 1)
    n_sent, n_tok, S = len(sentences), len(tokens), max_seq_length
    out: List[Window] = []
    tok_left = sent_left = i = 0
    while i < n_sent:
        tok_right = ends[i] + 1
        if tok_right - tok_left >= S - 1 or tok_right == n_tok:                     # :820
            ids = [cls_id] + tokens[tok_left:tok_right]
            lab = [IGNORE] + labels[tok_left:tok_right]
            ids, lab = ids[:S], lab[:S]                                               # :822-823 (sentence i may be cut)
            one_sentence = i == sent_left
            # mask the label of the window's last sentence (:843-849)
            last_bos = (ends[i - 1] + 1 if i > 0 else 0) - tok_left + 1
            if 0 <= last_bos < len(lab):
                lab[last_bos] = IGNORE
            first_sent = sent_left
            if one_sentence:
                tok_left = tok_right                                                  # :846
            else:
                tok_left = ends[i - 1] + 1                                            # :850 next window starts AT sentence i
            last_window = tok_right == n_tok
            present = range(first_sent, i + 1)
            if one_sentence or last_window:                                           # :855-857
                sent_left = i + 1
                i += 1
            else:                                                                     # :859 sentence i is re-read
                sent_left = i
            pad = S - len(ids)
            out.append(Window(ids + [pad_id] * pad, [1] * len(ids) + [0] * pad, lab + [IGNORE] * pad, present))
        else:
            i += 1
    return out


def collate(windows: Sequence[Window], device=None):
    """-> (input_ids, attention_mask, token_type_ids, labels) int64 [n_windows, S] (default_data_collator semantics)."""
    ids = torch.tensor([w.input_ids for w in windows], dtype=torch.long, device=device)
    mask = torch.tensor([w.attention_mask for w in windows], dtype=torch.long, device=device)
    labels = torch.tensor([w.labels for w in windows], dtype=torch.long, device=device)
    return ids, mask, torch.zeros_like(ids), labels


def synthetic_document(n_tokens: int, seed: int, bos_id: int = 30522, vocab_lo: int = 1000, vocab_hi: int = 30522,
                       sent_len=(10, 40), p_boundary: float = 0.15):
    """sentences of ~U[10,40] tokens, each prefixed by [BOS]; label 0 = 'B-EOP' (topic boundary
    after this sentence) with probability 0.15, else 1 = 'O'."""
    g = torch.Generator().manual_seed(seed)
    sents, labs, total = [], [], 0
    while total < n_tokens:
        n = int(torch.randint(sent_len[0], sent_len[1] + 1, (1,), generator=g))
        n = min(n, max(2, n_tokens - total))
        body = torch.randint(vocab_lo, vocab_hi, (n - 1,), generator=g).tolist()
        sents.append([bos_id] + body)
        labs.append(0 if float(torch.rand(1, generator=g)) < p_boundary else 1)
        total += n
    return sents, labs


def synthetic_segments(B: int, S: int, seed: int, pad_from=None):
    """SURVEY.md §8d config 3 inputs for PoNet: `segment_ids` as the reference driver builds them
    (`ponet_topic_segmentation.py:564-596`: 0 for [CLS], runs of ~U[8,40] tokens numbered 1..k, k+1 on padding) and the
    matching attention mask; row b is padded from position pad_from[b] on."""
    g = torch.Generator().manual_seed(seed)
    seg = torch.zeros(B, S, dtype=torch.long)
    mask = torch.ones(B, S, dtype=torch.long)
    for b in range(B):
        end = S if pad_from is None else pad_from[b]
        s, k = 1, 0
        while s < end:
            k += 1
            ln = int(torch.randint(8, 41, (1,), generator=g))
            seg[b, s:min(end, s + ln)] = k
            s += ln
        seg[b, end:] = k + 1
        mask[b, end:] = 0
    return seg, mask
