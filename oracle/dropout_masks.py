"""Numpy restatement of the B200 path's counter-based dropout masks (TEST INFRASTRUCTURE ONLY).

Not a restatement of the reference: the reference draws nn.Dropout masks from torch's Philox stream, and which elements
fall is not part of the algorithm.  The CUDA path instead hashes (step seed, site id, element index) — see
`spokennlp_b200/csrc/ptx.cuh` (`drop_hash`, `drop_seed`, `drop_pair`) — so that nothing is stored and the backward
regenerates the forward's mask.  This file repeats that hash on the host so the tests can hand the SAME masks to
`bert_oracle` (its `masks=` argument) and compare values and gradients under dropout.

Conventions.  The index enters the hash by addition: h = fin((idx + seed) * 0x9E3779B1), fin = xorshift 15, multiply
0x85EBCA6B, xorshift 13.
  hidden / embedding / classifier-input sites, tensor [rows, H]: one 32-bit hash per PAIR of consecutive elements (bits
    [0,15) -> even element, bits [16,31) -> odd element), pair index = (row * H + col) // 2; an element is kept iff its 15-bit
    lane >= round(p * 32768); kept elements are multiplied by 1/(1-p).
  attention-probability sites, tensor [B, heads, Sq, Sk]: one hash per FOUR consecutive keys (the attention kernels are
    instruction-issue bound; ptx.cuh: drop4_z), quad index = ((b*heads + h)*Sq + q) * ceil(Sk/4) + key//4, lane l = key % 4
    -> bits [8l, 8l+7); kept iff lane >= thr[l] where the four thresholds sum to t = round(512 p) (thr[l] = t//4 + (l < t%4));
    kept elements are multiplied by 512 / (512 - t), the inverse of the EFFECTIVE keep rate (p = 0.1: t = 51, rate 0.0996).
Site ids: layer*8 + {0: probabilities, 1: attention output dense, 2: FFN output dense, 3/4: cross-attention}, 0xE000
embeddings, 0xE001 classifier input (spokennlp_b200/engine.py: DropPlan).
"""
from __future__ import annotations

import numpy as np
import torch

M32 = np.uint64(0xFFFFFFFF)


def _hash(idx: np.ndarray, seed: np.ndarray) -> np.ndarray:
    h = (((idx.astype(np.uint64) + seed.astype(np.uint64)) & M32) * np.uint64(0x9E3779B1)) & M32
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x85EBCA6B)) & M32
    h ^= h >> np.uint64(13)
    return h


def site_seed(base_seed: int, site: int) -> np.ndarray:
    idx = np.array([(site * 0x632BE5AB + 0x7F4A7C15) & 0xFFFFFFFF], dtype=np.uint64)
    return _hash(idx, np.array([base_seed & 0xFFFFFFFF], dtype=np.uint64))


def _lanes(pairs: np.ndarray, base_seed: int, site: int, p: float):
    h = _hash(pairs, site_seed(base_seed, site))
    thr = np.uint64(int(np.float32(p) * np.float32(32768.0) + np.float32(0.5)))
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    lo = np.where((h & np.uint64(0x7FFF)) >= thr, scale, np.float32(0)).astype(np.float32)
    hi = np.where(((h >> np.uint64(16)) & np.uint64(0x7FFF)) >= thr, scale, np.float32(0)).astype(np.float32)
    return lo, hi


def hidden_mask(base_seed: int, site: int, p: float, rows: int, H: int) -> torch.Tensor:
    """[rows, H] multipliers (H even)."""
    pairs = np.arange(rows * H // 2, dtype=np.uint64)
    lo, hi = _lanes(pairs, base_seed, site, p)
    return torch.from_numpy(np.stack([lo, hi], axis=1).reshape(rows, H))


def prob_mask(base_seed: int, site: int, p: float, B: int, heads: int, Sq: int, Sk: int) -> torch.Tensor:
    """[B, heads, Sq, Sk] multipliers (quad format, see the module docstring)."""
    skq = (Sk + 3) // 4
    quads = np.arange(B * heads * Sq * skq, dtype=np.uint64)
    h = _hash(quads, site_seed(base_seed, site))
    t = int(np.float32(p) * np.float32(512.0) + np.float32(0.5))
    scale = np.float32(512.0) / (np.float32(512.0) - np.float32(t))
    lanes = []
    for l in range(4):
        thr = np.uint64(t // 4 + (1 if l < t % 4 else 0))
        lanes.append(np.where(((h >> np.uint64(8 * l)) & np.uint64(0x7F)) >= thr, scale, np.float32(0)).astype(np.float32))
    full = np.stack(lanes, axis=1).reshape(B, heads, Sq, 4 * skq)
    return torch.from_numpy(np.ascontiguousarray(full[..., :Sk]))


def bert_masks(base_seed: int, p_hidden: float, p_attn: float, n_layers: int, B: int, S: int, H: int, heads: int, head_site: bool = True):
    """The `masks` dict `bert_oracle.bert_model` / `topicseg_loss` take, for one training forward of the B200 engine."""
    m = {}
    if p_hidden > 0:
        m["emb"] = hidden_mask(base_seed, 0xE000, p_hidden, B * S, H).view(B, S, H)
        if head_site:
            m["head"] = hidden_mask(base_seed, 0xE001, p_hidden, B * S, H).view(B, S, H)
    for i in range(n_layers):
        pre = f"encoder.layer.{i}."
        if p_attn > 0:
            m[pre + "attention.probs"] = prob_mask(base_seed, i * 8 + 0, p_attn, B, heads, S, S)
        if p_hidden > 0:
            m[pre + "attention.out"] = hidden_mask(base_seed, i * 8 + 1, p_hidden, B * S, H).view(B, S, H)
            m[pre + "ffn_out"] = hidden_mask(base_seed, i * 8 + 2, p_hidden, B * S, H).view(B, S, H)
    return m
