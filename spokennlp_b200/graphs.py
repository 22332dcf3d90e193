"""Replay a whole step of the drop-in modules from one CUDA graph.

The cross-modal path of mmvts runs a few hundred launches over a few hundred token rows per step (N = 300 clips):
eager, its step time is Python + launch overhead (6 ms whether the batch holds 1 or 16 videos, profiles/config_sweeps_r02.jsonl).
`GraphedStep` captures `fn(*inputs)` — forward, loss and backward of any of the drop-in modules, optionally a capturable optimizer —
once and replays it; inputs are copied into the captured buffers, outputs and `.grad`s live in the graph's memory pool and are
overwritten by the next replay.

    step = GraphedStep(lambda t, v, a, m: loss_fn(encoder(m, *projector(t, v, a))), (tfeat, vfeat, afeat, mask))
    loss = step(tfeat2, vfeat2, afeat2, mask2)          # same shapes / dtypes; gradients in p.grad

What makes the modules capture-safe: no host synchronisation or pageable host copies on their paths, dropout seeds that advance on
the device (modeling_cross._Packed._drop_plan), and the fp16 operand mirror refreshed inside the graph (`_flat`).  `fn` must not
set `.grad = None` on parameters whose gradients the caller reads after a replay only if it does so on every call (the captured
backward then writes each gradient into the same graph-owned tensor).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from .lib import B200Error

Tensor = torch.Tensor


class GraphedStep:
    def __init__(self, fn: Callable, example_inputs: Sequence[Tensor], warmup: int = 3):
        if not example_inputs or any(not t.is_cuda for t in example_inputs):
            raise B200Error("GraphedStep: inputs must be CUDA tensors (no CPU fallback)")
        self.fn = fn
        self.static_in = [t.detach().clone().requires_grad_(t.requires_grad) for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up off the capture stream: lazy initialisation (packed parameter
            for _ in range(max(1, warmup)):                 # buffers, TMA descriptors, kernel attributes, dropout seeds) happens here
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs: Tensor):
        if len(inputs) != len(self.static_in):
            raise B200Error(f"GraphedStep: expected {len(self.static_in)} inputs, got {len(inputs)}")
        for dst, src in zip(self.static_in, inputs):
            if src.shape != dst.shape or src.dtype != dst.dtype:
                raise B200Error(f"GraphedStep: input {tuple(src.shape)} {src.dtype} does not match the captured {tuple(dst.shape)} {dst.dtype}")
            if src.data_ptr() != dst.data_ptr():
                dst.detach().copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
