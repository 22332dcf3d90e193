"""Data-parallel fine-tuning step for the topic-segmentation path (BASELINE config 2).

What one step replaces in the reference: `BertWithDAForSentenceLabelingTopicSegmentation.forward`
(bert_for_ts.py:35-113, `ts_score_predictor == "lt"` path: encoder -> Linear(H,2) -> CE over [BOS] rows),
`loss.backward()`, DDP's gradient allreduce, `clip_grad_norm_(1.0)`, `AdamW.step`, linear-decay schedule
(HF Trainer defaults).  HF `Trainer` itself cannot be constructed here (no `accelerate`).

One process per GPU.  The only inter-GPU exchange is the gradient allreduce (NCCL over NVLink), issued per encoder
layer as soon as that layer's weight gradients are complete so that it overlaps the rest of the backward.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.distributed as dist
from torch import nn

from . import ops
from .engine import EMB_NAMES, DropPlan, EncoderEngine, FlatParams, layer_param_names
from .modeling_bert import BertModel


def gradient_buckets(layer_first_offsets, numel: int):
    """Contiguous slices of the flat gradient buffer, in the order their gradients become final during backward:
    [layer L-1 + head], ..., [layer 0], [embeddings].  One allreduce per slice."""
    offs = list(layer_first_offsets) + [numel]
    layers = [(offs[i], offs[i + 1]) for i in range(len(layer_first_offsets))]
    return list(reversed(layers)) + [(0, offs[0])]


def allreduce_bucket(flat_grad: torch.Tensor, lo: int, hi: int, group=None, async_op: bool = True):
    """Sum-allreduce one bucket in place (NCCL on GPUs; gloo in the CPU tests).  Averaging is folded into the optimizer's
    gradient multiplier (1 / world size), so there is no second pass over the gradients."""
    if hi <= lo:
        return None
    return dist.all_reduce(flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class TopicSegModel(nn.Module):
    """Encoder + token-classification head with the reference wrapper's parameter names
    (`bert.*`, `loss_calculator.classifier.*`: bert_for_ts.py:23, loss_calculator.py:17)."""

    def __init__(self, config, num_labels: int = 2):
        super().__init__()
        self.config = config
        self.bert = BertModel(config, add_pooling_layer=False)
        self.loss_calculator = nn.Module()
        self.loss_calculator.classifier = nn.Linear(config.hidden_size, num_labels)
        nn.init.normal_(self.loss_calculator.classifier.weight, std=config.initializer_range)
        nn.init.zeros_(self.loss_calculator.classifier.bias)


class DataParallelTrainer:
    def __init__(self, model: TopicSegModel, *, lr: float = 5e-5, total_steps: int = 1000, max_grad_norm: float = 1.0,
                 weight_decay: float = 0.0, loss_scale: float = 32768.0, device: Optional[torch.device] = None,
                 dropout: bool = True, seed: int = 0, dynamic_loss_scale: bool = True, scale_growth_interval: int = 2000):
        """`dropout=True` honours the config's hidden_dropout_prob / attention_probs_dropout_prob (the reference trains with
        both at 0.1); `seed` is the base of the per-step dropout seeds.  `loss_scale` is the INITIAL scale of the fp16
        activation gradients; with `dynamic_loss_scale` a step whose gradient norm is not finite is skipped, halves the scale
        and is counted (`skipped_steps()`), and the scale doubles again after `scale_growth_interval` finite steps — all on
        the device, so a replayed CUDA graph adapts too."""
        self.model = model
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        cfg = model.config
        own = dict(model.named_parameters())
        names = ["bert." + n for n in EMB_NAMES]
        self.layer_first = []
        for i in range(cfg.num_hidden_layers):
            self.layer_first.append("bert." + layer_param_names(i)[0])
            names += ["bert." + n for n in layer_param_names(i)]
        names += ["loss_calculator.classifier.weight", "loss_calculator.classifier.bias"]
        flat = FlatParams([(n, own[n]) for n in names], self.device)
        # the engine addresses parameters by their BertModel-relative names
        flat_alias = _Aliased(flat, "bert.")
        self.flat = flat
        self.engine = EncoderEngine(flat_alias, cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size,
                                    cfg.num_hidden_layers, float(cfg.layer_norm_eps),
                                    pad_id=model.bert.embeddings.word_embeddings.padding_idx)
        model.bert._engine = self.engine
        flat.ensure_grad()
        self.m = torch.zeros_like(flat.flat32)
        self.v = torch.zeros_like(flat.flat32)
        self.lr, self.total_steps, self.max_grad_norm, self.wd = lr, total_steps, max_grad_norm, weight_decay
        self.step_idx = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        # B200_COMM_CTAS=n (default 4; 0 = off): overlapped gradient buckets run on a second NCCL communicator capped at n CTAs,
        # and the persistent compute kernels of the backward (one CTA per SM, static work split) leave that many SMs free —
        # a compute CTA that has to wait for a communication kernel to vacate its SM delays its whole kernel.  The last
        # bucket, which nothing overlaps, still goes through the default communicator at full width.  Measured on 2 GPUs
        # (profiles/bench_r01j_n2_commctas{0,4}.json, same box): 4029 -> 4161 seq/s with n = 4; n = 8 and n <= 2 were
        # slower.  Round 2 (profiles/scale_r02_*): 4301 -> 4374 seq/s on 2 GPUs inside the captured graph; validated on 4 / 8 GPUs.
        self.comm_ctas, self.bg_group, self._sms = 0, None, 0
        if self.world > 1 and self.device.type == "cuda" and dist.get_backend() == "nccl":
            want = int(os.environ.get("B200_COMM_CTAS", "4"))        # default on since round 2 (2 GPUs: 14.88 -> 14.63 ms/step; 0 = off)
            if want > 0:
                try:
                    opts = dist.ProcessGroupNCCL.Options()
                    opts.config.max_ctas = want
                    opts.config.min_ctas = 1
                    self.bg_group = dist.new_group(pg_options=opts)     # collective call: every rank constructs its trainer
                    self.comm_ctas = want
                    self._sms = torch.cuda.get_device_properties(self.device).multi_processor_count
                except Exception as e:  # noqa: BLE001  (older NCCL / torch without per-communicator config)
                    import warnings
                    warnings.warn(f"capped background communicator unavailable ({type(e).__name__}: {e}); using the default one")
                    self.bg_group, self.comm_ctas = None, 0
        self.scale = torch.tensor([loss_scale, 1.0 / loss_scale], dtype=torch.float32, device=self.device)
        self.dynamic_loss_scale, self.scale_growth_interval = bool(dynamic_loss_scale), int(scale_growth_interval)
        self.scale_state = torch.zeros(2, dtype=torch.float32, device=self.device)     # {consecutive finite steps, skipped steps}
        self.stats = torch.zeros(2, dtype=torch.float32, device=self.device)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.coef = torch.zeros(3, dtype=torch.float32, device=self.device)
        self.num_labels = model.loss_calculator.classifier.weight.shape[0]
        # gradient buckets: [embeddings | layer 0 | ... | layer L-1 + head], contiguous slices of the flat buffer
        buckets = gradient_buckets([flat.offsets[n] for n in self.layer_first], flat.numel)
        self.layer_slices = list(reversed(buckets[:-1]))      # indexed by layer
        self.emb_slice = buckets[-1]
        self._works = []
        self.hyper = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._graph = None
        self._static = None
        self.kernels_per_step = 0
        # AdamW clears the gradient buffer while it streams through it (0.44 GB less traffic per step than a separate fill);
        # callers that want to inspect gradients AFTER optimizer_step set this to False
        self.fused_zero_grad = True
        self._grads_clean = True            # ensure_grad() zero-filled the buffer
        # dropout (HF config probabilities; TopicSegModel's classifier dropout = hidden_dropout_prob, bert_for_ts.py:66)
        self.p_hidden = float(getattr(cfg, "hidden_dropout_prob", 0.0)) if dropout else 0.0
        self.p_attn = float(getattr(cfg, "attention_probs_dropout_prob", 0.0)) if dropout else 0.0
        self.seed = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.base_seed = int(seed) & 0x7FFFFFFF
        self.drop = DropPlan(self.seed, self.p_hidden, self.p_attn) if (self.p_hidden > 0 or self.p_attn > 0) else None

    # ------------------------------------------------------------------------------------------------------------
    def _allreduce_slice(self, lo: int, hi: int, group=None) -> None:
        if self.world > 1 and hi > lo:
            self._works.append(allreduce_bucket(self.flat.grad32, lo, hi, group=group))

    def _after_layer(self, i: int) -> None:
        if i >= 0:
            self._allreduce_slice(*self.layer_slices[i], group=self.bg_group)      # overlaps the rest of the backward
        else:
            self._allreduce_slice(*self.emb_slice)                                 # exposed: full-width communicator

    def _reserve_comm_sms(self, on: bool) -> None:
        if self.comm_ctas > 0:
            from . import lib as _lib
            reserve = int(os.environ.get("B200_COMM_RESERVE", str(self.comm_ctas)))
            _lib.load().b200_set_sm_limit(self._sms - reserve if (on and reserve > 0) else 0)

    def forward_backward(self, input_ids, attention_mask, token_type_ids, labels, pack: bool = False) -> torch.Tensor:
        """Enqueue forward + backward for one local batch; returns the device scalar loss (no sync).
        `pack=True`: the right-padded windows are packed — only the valid tokens go through the
        encoder, attention walks the sequences by cu_seqlens — at the price of ONE 8-byte host read (the packed row count sizes
        the activation buffers), so a packed step is launched eagerly, not replayed from a CUDA graph."""
        eng, flat = self.engine, self.flat
        B, S = input_ids.shape
        if not self._grads_clean:           # the fused optimizer step leaves the gradient buffer zeroed
            flat.grad32.zero_()
        self._grads_clean = False
        key_bias = kv_len = rows = None
        pos = None  # arange(S) per row
        ids = input_ids.contiguous().view(-1)
        tt = token_type_ids.contiguous().view(-1) if token_type_ids is not None else None
        if pack and attention_mask is not None:
            rows = ops.compact_rows(attention_mask, 0)
            ids = ops.gather_i64(ids, rows)
            tt = ops.gather_i64(tt, rows) if tt is not None else None
            labels = ops.gather_i64(labels, rows)
            pos = rows.rank.to(torch.int64)
        elif attention_mask is not None:
            key_bias, kv_len = ops.mask_to_bias(attention_mask)
        x16, _, saved, _, _ = eng.forward(ids, tt, pos, None, key_bias, kv_len, B, S, save=True, drop=self.drop, pack=rows)
        drop_head = self.drop.at(DropPlan.HEAD, self.p_hidden) if self.drop is not None else None
        W32 = flat.view32("loss_calculator.classifier.weight")
        b32 = flat.view32("loss_calculator.classifier.bias")
        logits = ops.cls_head_fwd(x16, W32, b32, drop=drop_head)
        lab = labels.contiguous().view(-1)
        self.stats.zero_()
        ops.ce_stats(logits, lab, self.stats)
        dh = torch.empty_like(x16)
        ops.cls_head_bwd(x16, logits, lab, self.stats, W32, dh, flat.viewg("loss_calculator.classifier.weight"),
                         flat.viewg("loss_calculator.classifier.bias"), scale=self.scale[0:1], drop=drop_head)
        self._works = []
        # the head's gradients live in the last bucket, which is reduced right after the top layer's backward
        self._reserve_comm_sms(True)
        try:
            eng.backward(saved, dh, self.scale[1:2], after_layer=self._after_layer)
        finally:
            self._reserve_comm_sms(False)
        self.last_logits = logits
        return self.stats

    def _lr(self) -> float:
        return self.lr * max(0.0, 1.0 - (self.step_idx - 1) / max(1, self.total_steps))     # HF linear schedule, 0 warm-up

    def optimizer_step(self, *, advance: bool = True) -> None:
        """Global-norm clip + fused AdamW over the flat buffers.  Hyper-parameters are read from device memory
        (`self.hyper`, written by `_push_hyper`) so the same launches can live inside a captured CUDA graph."""
        flat = self.flat
        if advance:
            self.step_idx += 1
            self._push_hyper()
        self.sumsq.zero_()
        if len(self._works) > 1:
            # the trailing embeddings bucket is the one exposed collective of the step (95 MB, ~0.4 ms on 8 GPUs): the norm of
            # everything else (78 % of the gradient bytes) is summed while it is still in flight
            for w in self._works[:-1]:
                w.wait()
            lo = self.emb_slice[1]
            ops.grad_sumsq(flat.grad32[lo:], self.sumsq)
            self._works[-1].wait()
            ops.grad_sumsq(flat.grad32[:lo], self.sumsq)
        else:
            for w in self._works:
                w.wait()
            ops.grad_sumsq(flat.grad32, self.sumsq)
        self._works = []
        if self.dynamic_loss_scale:
            ops.clip_coef_scaled(self.sumsq, self.coef, self.max_grad_norm, 1.0 / self.world, self.scale, self.scale_state,
                                 growth_interval=self.scale_growth_interval)
        else:
            ops.clip_coef(self.sumsq, self.coef, self.max_grad_norm, 1.0 / self.world)
        ops.adamw_step_dev(flat.flat32, flat.grad32, self.m, self.v, flat.flat16, self.hyper, self.coef, zero_grad=self.fused_zero_grad)
        self._grads_clean = self.fused_zero_grad
        flat.version = flat.cur_version()       # the fused step refreshed the fp16 mirror itself

    def _push_hyper(self) -> None:
        """Per-step device-side state that a replayed graph must see change: learning rate and bias corrections."""
        ops.set_hyper(self.hyper, lr=self._lr(), weight_decay=self.wd, step=max(1, self.step_idx))

    def step_seed(self, step_idx: int) -> int:
        """Base dropout seed of training step `step_idx` on this rank (ranks draw different masks, as DDP replicas do)."""
        return (self.base_seed * 1000003 + step_idx * 7919 + self.rank * 104729) & 0x7FFFFFFF

    def _push_seed(self) -> None:
        """Written OUTSIDE the captured graph, before the step's forward: the kernels read the seed from device memory."""
        if self.drop is not None:
            self.seed.fill_(self.step_seed(self.step_idx))

    def release_graph(self) -> None:
        """Drop the captured step (must precede tearing down the process group: the graph holds NCCL work)."""
        self._graph = None
        torch.cuda.synchronize()

    def step(self, input_ids, attention_mask, token_type_ids, labels, pack: bool = False) -> torch.Tensor:
        if self._graph is not None and not pack:
            return self._replay(input_ids, attention_mask, token_type_ids, labels)
        self._push_seed()
        stats = self.forward_backward(input_ids, attention_mask, token_type_ids, labels, pack=pack)
        self.optimizer_step()
        return stats

    # ---- CUDA graph: one launch per step instead of ~270 ---------------------------------------------------------------
    def capture(self, input_ids, attention_mask, token_type_ids, labels, warmup: int = 2) -> bool:
        """Capture forward + backward (+ per-layer allreduce) + optimizer of one step for this batch SHAPE into a CUDA
        graph.  Later `step()` calls copy their batch into the static input buffers and replay.  The learning-rate
        schedule keeps advancing because AdamW reads its hyper-parameters from device memory.  Returns False (and stays
        eager) if capture is not possible in this process."""
        self._static = [t.clone() if t is not None else None for t in (input_ids, attention_mask, token_type_ids, labels)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        # The warm-up steps are REAL steps (they must exercise every kernel, allocation and collective the capture will see),
        # so everything they change is put back afterwards: parameters, their fp16 mirror, the moments, the loss scale, the
        # step counter and learning-rate schedule — a captured run starts from the same state as an eager one.
        flat = self.flat
        snap = [t.clone() for t in (flat.flat32, flat.flat16, self.m, self.v, self.scale, self.scale_state)]
        step0, clean0 = self.step_idx, self._grads_clean
        try:
            with torch.cuda.stream(side):
                for _ in range(warmup):                       # warm-up off the capture: lazy init, allocator pools, NCCL
                    self._push_seed()
                    self.forward_backward(*self._static)
                    self.optimizer_step()
                for dst, src in zip((flat.flat32, flat.flat16, self.m, self.v, self.scale, self.scale_state), snap):
                    dst.copy_(src)
                if not self._grads_clean:
                    flat.grad32.zero_()
                    self._grads_clean = True
            self.step_idx = step0
            flat.version = flat.cur_version()                 # the fp16 mirror was restored together with the fp32 buffer
            del snap, clean0
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            from . import lib as _lib
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph):
                self.forward_backward(*self._static)
                self.optimizer_step(advance=False)
            self.kernels_per_step = _lib.launch_count() - n0
            self._graph = graph
            return True
        except Exception as e:  # noqa: BLE001  (capture is an optimisation; the eager path is the same kernels)
            import warnings
            warnings.warn(f"CUDA graph capture of the training step failed ({type(e).__name__}: {e}); staying eager")
            self._graph = None
            torch.cuda.synchronize()
            return False

    def _replay(self, input_ids, attention_mask, token_type_ids, labels) -> torch.Tensor:
        for dst, src in zip(self._static, (input_ids, attention_mask, token_type_ids, labels)):
            if dst is not None and src is not dst:
                dst.copy_(src, non_blocking=True)
        self._push_seed()
        self.step_idx += 1
        self._push_hyper()
        self._graph.replay()
        return self.stats

    def loss_value(self) -> float:
        s = self.stats.tolist()     # device -> host read
        return s[0] / max(s[1], 1e-30)

    def skipped_steps(self) -> int:
        """Optimizer steps skipped so far because the gradient norm was not finite (device -> host read)."""
        return int(self.scale_state[1].item())

    def loss_scale_value(self) -> float:
        return float(self.scale[0].item())

    # ---- end-to-end step: pinned host batch in, loss out ------------------------------------------------------------------
    def step_from_host(self, input_ids, attention_mask, token_type_ids, labels) -> Optional[float]:
        """One training step on a batch that lives in PINNED host memory: host->device copies of the four tensors, the step,
        and a device->host read of its loss statistics, all enqueued on the current stream.  The loss of EVERY step is read
        back, one step late: the call returns the loss of the previous step (None on the first call) after waiting only for
        that step's copy event, so the host never idles the GPU between steps.  `drain()` returns the last step's loss."""
        if self._static is not None and self._graph is not None:
            for dst, src in zip(self._static, (input_ids, attention_mask, token_type_ids, labels)):
                if dst is not None:
                    dst.copy_(src, non_blocking=True)           # straight into the captured graph's input buffers
            batch = self._static
        else:
            batch = [t.cuda(non_blocking=True) if t is not None else None for t in (input_ids, attention_mask, token_type_ids, labels)]
        self.step(*batch)
        if not hasattr(self, "_host_stats"):
            self._host_stats = [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._host_events = [torch.cuda.Event(), torch.cuda.Event()]
            self._host_pending = [False, False]
            self._host_k = 0
        k = self._host_k & 1
        self._host_stats[k].copy_(self.stats, non_blocking=True)
        self._host_events[k].record()
        self._host_pending[k] = True
        self._host_k += 1
        return self._read_host_loss(k ^ 1)

    def _read_host_loss(self, k: int) -> Optional[float]:
        if not getattr(self, "_host_pending", [False, False])[k]:
            return None
        self._host_events[k].synchronize()
        self._host_pending[k] = False
        s = self._host_stats[k]
        return float(s[0]) / max(float(s[1]), 1e-30)

    def drain(self) -> Optional[float]:
        """Loss of the most recent `step_from_host` call (waits for it)."""
        if not hasattr(self, "_host_k") or self._host_k == 0:
            return None
        return self._read_host_loss((self._host_k - 1) & 1)


class _Aliased:
    """FlatParams view that resolves BertModel-relative names inside a larger (prefixed) flat buffer."""

    def __init__(self, flat: FlatParams, prefix: str):
        self._f, self._p = flat, prefix

    def view32(self, name, extra=()):
        return self._f.view32(self._p + name, tuple(self._p + e for e in extra))

    def view16(self, name, extra=()):
        return self._f.view16(self._p + name, tuple(self._p + e for e in extra))

    def viewg(self, name, extra=()):
        return self._f.viewg(self._p + name, tuple(self._p + e for e in extra))

    def sync_half(self, force: bool = False):
        return self._f.sync_half(force)

    def intact(self):
        return self._f.intact()

    @property
    def names(self):
        return [n[len(self._p):] for n in self._f.names if n.startswith(self._p)]

    @property
    def params(self):
        return {n[len(self._p):]: p for n, p in self._f.params.items() if n.startswith(self._p)}

    @property
    def flat32(self):
        return self._f.flat32

    @property
    def flat16(self):
        return self._f.flat16

    @property
    def grad32(self):
        return self._f.grad32

    @grad32.setter
    def grad32(self, v):
        self._f.grad32 = v
