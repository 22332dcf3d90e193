#!/usr/bin/env python
"""Headline benchmark: 512-token sequences/second of BERT-base topic-segmentation fine-tuning on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's CPU path on the host cores)

A "step" = one fine-tuning step on one synthetic batch of 32 x 512-token windows per GPU (BASELINE config 2,
SURVEY.md §8d): embeddings -> 12 encoder layers -> token-cls head -> CE -> full backward -> gradient allreduce ->
clip(1.0) -> AdamW.  value = whole-job sequences/s with the batch already resident in HBM; e2e = same through
`DataParallelTrainer.step_from_host` with pinned-host inputs copied every step and the loss read back.  Both loops start from an idle
part (1 s) plus their warm-up steps; `sustained` repeats the resident-batch step for >= 3 s under its own clock sample.  `roofline` /
`kernels` time each hot kernel alone as the best of five short bursts (the way the burst peaks in MEASURED_PEAKS.json were taken),
before the sustained run.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEQ, BATCH, VOCAB = 512, 32, 30523
FLOP_PER_SEQ = 289_910_292_480            # fwd+bwd, SURVEY.md §8d / BASELINE.md §3
CFG = dict(hidden_size=768, num_attention_heads=12, intermediate_size=3072, num_hidden_layers=12, vocab_size=VOCAB,
           max_position_embeddings=512, type_vocab_size=2)


def synth_batch(torch, batch, seq, seed, padded=False):
    """SURVEY.md §8d synthetic inputs for config 2."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, 30522, (batch, seq), generator=g)
    ids[:, 0] = 101
    bos = torch.arange(1, seq, 20)
    ids[:, bos] = 30522
    mask = torch.ones(batch, seq, dtype=torch.long)
    if padded:
        lens = torch.randint(seq // 2, seq + 1, (batch,), generator=g)
        mask = (torch.arange(seq)[None, :] < lens[:, None]).long()
    tt = torch.zeros(batch, seq, dtype=torch.long)
    labels = torch.full((batch, seq), -100, dtype=torch.long)
    labels[:, bos] = (torch.rand(batch, len(bos), generator=g) < 0.85).long()
    labels = torch.where(mask.bool(), labels, torch.full_like(labels, -100))
    return ids, mask, tt, labels


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0 = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def begin(self):
        """The timed region starts now: nvidia-smi takes a few hundred ms to produce its first sample, so the sampler is started
        well before a 0.3 s region and only the samples taken from here on are reported."""
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.time()
        time.sleep(0.05)                       # (a sample is printed a little after it was taken)
        self.proc.terminate()
        inside = [r for t, r in self.rows if self.t0 is None or self.t0 <= t <= t1 + 0.05]
        sm, mx, reasons, pw = [], [], set(), []
        for r in (inside or [r for _, r in self.rows[-3:]]):
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference (CPU) arm
def cpu_reference_step_fn(torch, batch, dropout=0.1):
    """The reference's own implementation of the path on host cores: HuggingFace `BertModel` (eager, fp32) — the class
    bert_for_ts.py:7 imports — + Linear(768,2) + CE + backward + clip + AdamW.  Falls back to the oracle port."""
    ids, mask, tt, labels = synth_batch(torch, batch, SEQ, 1234)
    try:
        from transformers import BertConfig, BertModel
        cfg = BertConfig(attn_implementation="eager", hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout, **CFG)
        torch.manual_seed(0)
        bert = BertModel(cfg, add_pooling_layer=False).train()
        head = torch.nn.Sequential(torch.nn.Dropout(dropout), torch.nn.Linear(768, 2)).train()      # bert_for_ts.py:66-67
        params = list(bert.parameters()) + list(head.parameters())
        opt = torch.optim.AdamW(params, lr=5e-5, weight_decay=0.0)

        def step():
            opt.zero_grad(set_to_none=True)
            h = bert(ids, attention_mask=mask, token_type_ids=tt, return_dict=False)[0]
            loss = torch.nn.functional.cross_entropy(head(h).view(-1, 2), labels.view(-1))
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            return float(loss.detach())
        return step, "reference"
    except Exception:
        from oracle import bert_oracle as O
        ocfg = O.OracleConfig(**CFG)
        sd = {k: v.requires_grad_(not k.startswith("pooler")) for k, v in O.random_state_dict(ocfg, 0).items()}
        w = (torch.randn(2, 768) * 0.02).requires_grad_(True)
        b = torch.zeros(2, requires_grad=True)
        params = [p for p in sd.values() if p.requires_grad] + [w, b]
        opt = torch.optim.AdamW(params, lr=5e-5, weight_decay=0.0)

        def step():
            opt.zero_grad(set_to_none=True)
            loss, _ = O.topicseg_loss(sd, ocfg, w, b, ids, mask, tt, labels)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            return float(loss.detach())
        return step, "port"


def time_cpu_reference(torch, steps, warmup, batch, dropout=0.1):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = cpu_reference_step_fn(torch, batch, dropout)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, steps)
    return {"value": batch / dt, "unit": "seq/s", "cores": cores, "kind": kind, "ms_per_step": dt * 1e3,
            "sample": f"{steps} fwd+bwd+AdamW steps of [{batch},{SEQ}] BERT-base fp32 on {cores} host threads "
                      f"(throughput is linear in batch)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    steps, warmup = max(1, min(args.steps, 20)), max(1, min(args.warmup, 2))
    r = time_cpu_reference(torch, steps, warmup, batch=1, dropout=args.dropout)
    line = {"impl": "reference", "metric": "512-tok seq/sec BERT-base topic-seg fine-tune", "value": r["value"], "unit": "seq/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "emnlp2023-topic_segmentation BERT-base fine-tune, 512-tok windows (CPU reference path; "
                                   "each step is a bounded [1,512] sample of the bsz-32 workload)", "seq_len": SEQ,
                       "global_batch": 1, "parallelism": "cpu", "dropout": args.dropout},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def time_kernel(torch, fn, iters=10, warmup=3, reps=5):
    """Seconds per launch of a kernel timed alone: the BEST of `reps` bursts of `iters` back-to-back launches (CUDA events on the
    launch stream).  MEASURED_PEAKS.json's burst peaks are best-of-10 figures of a kernel timed alone; an average over one long burst
    instead folds in the part's power state — the same kernel reads 5-10 % slower after a few tens of milliseconds at the cap."""
    for _ in range(warmup):
        fn()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e-3)
    return best


def kernel_rooflines(torch, ops, lib, peaks, dropout):
    """Live per-kernel numbers on the bench shapes (each kernel timed alone with CUDA events on the launch stream;
    operands >> L2 for the big GEMMs)."""
    from spokennlp_b200 import blocks
    M, H, I = BATCH * SEQ, 768, 3072
    dev = "cuda"
    f16 = torch.float16
    x = torch.randn(M, H, device=dev, dtype=f16)
    w1 = torch.randn(I, H, device=dev, dtype=f16) * 0.02
    b1 = torch.zeros(I, device=dev)
    h = torch.empty(M, I, device=dev, dtype=f16)
    z = torch.empty(M, I, device=dev, dtype=f16)
    out = {}
    t = time_kernel(torch, lambda: ops.gemm(x, w1, h, epilogue=ops.EPI_BIAS_GELU, bias=b1, out2=z))
    flops = 2.0 * M * H * I
    out["gemm_ffn_up_gelu"] = {"bound": "tensor", "achieved": flops / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3}
    try:        # yardstick only (never on the product path): the library's plain fp16 GEMM of the same shape, no epilogue
        tl = time_kernel(torch, lambda: torch.matmul(x, w1.t(), out=h))
        out["gemm_ffn_up_gelu"]["cublas_plain_same_shape_tflops"] = flops / tl / 1e12
    except Exception as e:  # noqa: BLE001
        out["gemm_ffn_up_gelu"]["cublas_plain_same_shape_tflops"] = None
        out["gemm_ffn_up_gelu"]["cublas_error"] = f"{type(e).__name__}: {e}"[:200]
    # the remaining kernels are timed AS THE STEP RUNS THEM: dropout 0.1 where the training step has it, the fused entry points
    # of the default schedule (in-place residual accumulate, row statistic ready), every buffer allocated outside the timed calls
    seed = torch.tensor([7], dtype=torch.int32, device=dev)
    p = float(dropout)
    drop_h = ops.Dropout(seed, 2, p) if p > 0 else None
    drop_a = ops.Dropout(seed, 0, p) if p > 0 else None
    w2 = torch.randn(H, I, device=dev, dtype=f16) * 0.02
    pre = torch.randn(M, H, device=dev, dtype=torch.float32)
    bo = torch.zeros(H, device=dev)
    t = time_kernel(torch, lambda: ops.gemm_resadd(h, w2, pre, bo, drop=drop_h))
    out["gemm_ffn_down_res"] = {"bound": "tensor", "achieved": flops / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3, "dropout": p}
    ctx = torch.randn(M, H, device=dev, dtype=f16)
    wo = torch.randn(H, H, device=dev, dtype=f16) * 0.02
    t = time_kernel(torch, lambda: ops.gemm_resadd(ctx, wo, pre, bo, drop=drop_h))
    out["gemm_out_proj_res"] = {"bound": "tensor", "achieved": 2.0 * M * H * H / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3, "dropout": p}
    gw = torch.zeros(I, H, device=dev)
    t = time_kernel(torch, lambda: ops.gemm(h, x, gw, a_layout=1, b_layout=1, epilogue=ops.EPI_ATOMIC, k_splits=ops.wgrad_splits(I, H, M)))
    out["gemm_wgrad_ffn_up"] = {"bound": "tensor", "achieved": flops / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3}
    dz, db1 = torch.empty(M, I, device=dev, dtype=f16), torch.zeros(I, device=dev)
    t = time_kernel(torch, lambda: ops.gemm_dgelu_colsum(x, w2, z, dz, db1))
    out["gemm_dgrad_ffn_down_dgelu_colsum"] = {"bound": "tensor", "achieved": flops / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3}
    qkv = torch.randn(M, 3 * H, device=dev, dtype=f16)
    lse = torch.empty(BATCH, 12, SEQ, device=dev)
    t = time_kernel(torch, lambda: ops.attn_fwd(qkv, qkv, ctx, BATCH, 12, SEQ, SEQ, q_col0=0, k_col0=H, v_col0=2 * H, lse2=lse, drop=drop_a))
    aflops = 4.0 * BATCH * SEQ * SEQ * H
    out["attn_fwd"] = {"bound": "tensor", "achieved": aflops / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3, "dropout": p}
    dqkv = torch.empty_like(qkv)
    ws = ops.attn_bwd_workspace(BATCH, 12, SEQ, dev)
    dctx = torch.randn(M, H, device=dev, dtype=f16) * 0.1
    ops.gemm_dgrad_delta(dctx, wo, ctx, dctx.clone(), ws, BATCH, 12, SEQ)          # fills the row statistic once
    t = time_kernel(torch, lambda: ops.attn_bwd(qkv, qkv, dctx, ctx, lse, dqkv, dqkv, ws, BATCH, 12, SEQ, SEQ, q_col0=0, k_col0=H,
                                                v_col0=2 * H, dq_col0=0, dk_col0=H, dv_col0=2 * H, drop=drop_a, delta_ready=True,
                                                dq_half="dq16" in blocks.Experimental.active()))
    out["attn_bwd"] = {"bound": "tensor", "achieved": 2 * aflops / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3, "dropout": p,
                       "includes": ("memset of the dQ columns + attn_bwd3_kernel (dQ as fp16 TMA reduce-adds" if "dq16" in blocks.Experimental.active()
                                    else "dq_acc memset + attn_bwd3_kernel + dq_cast_kernel (") + "; row statistic from the out-proj dgrad epilogue)"}
    g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    y = torch.empty(M, H, device=dev, dtype=f16)
    y32 = torch.empty(M, H, device=dev)
    mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
    t = time_kernel(torch, lambda: ops.layernorm_fwd(pre, g, b, 1e-12, y=y, y32=y32, mean=mean, rstd=rstd))
    out["layernorm_fwd"] = {"bound": "hbm", "achieved": M * H * (4 + 2 + 4) / t / 1e9, "unit": "GB/s", "ms": t * 1e3,
                            "bytes": "read fp32 pre-LN sum, write fp16 operand copy + fp32 residual copy (7680 B/row)"}
    dxl, dxd = torch.empty(M, H, device=dev, dtype=f16), torch.empty(M, H, device=dev, dtype=f16)
    dgl, dbl, dbias = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
    if drop_h is not None:
        t = time_kernel(torch, lambda: ops.layernorm_bwd(y, pre, mean, rstd, g, dxl, dgl, dbl, dbias=dbias, dx_drop=dxd, drop=drop_h))
        nbytes = M * H * (2 + 4 + 2 + 2)
    else:
        t = time_kernel(torch, lambda: ops.layernorm_bwd(y, pre, mean, rstd, g, dxl, dgl, dbl, dbias=dbias))
        nbytes = M * H * (2 + 4 + 2)
    out["layernorm_bwd"] = {"bound": "hbm", "achieved": nbytes / t / 1e9, "unit": "GB/s", "ms": t * 1e3, "dropout": p}
    W, bcls = torch.randn(2, H, device=dev) * 0.02, torch.zeros(2, device=dev)
    logits = torch.empty(M, 2, device=dev)
    import ctypes as C
    so, st = lib.load(), C.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda tns: C.c_void_p(tns.data_ptr())
    t = time_kernel(torch, lambda: so.b200_cls_head_fwd(vp(y), vp(W), vp(bcls), vp(logits), None, M, H, 2, st), iters=50)
    out["cls_head_fwd"] = {"bound": "hbm", "achieved": (M * H * 2 + M * 8) / t / 1e9, "unit": "GB/s", "ms": t * 1e3}
    for v in out.values():
        peak = peaks["bf16_tflops"] if v["bound"] == "tensor" else peaks["hbm_gbs"]
        v["peak"], v["frac"] = peak, v["achieved"] / peak
    return out


def arm_watchdog(rank, seconds):
    """A collective that never completes must not hang the caller for ever: after `seconds` without the bench finishing,
    rank 0 prints a JSON line that says so and every rank leaves with exit code 3."""
    def fire():
        if rank == 0:
            print(json.dumps({"metric": "512-tok seq/sec BERT-base topic-seg fine-tune", "value": None, "unit": "seq/s",
                              "error": f"bench.py watchdog: no result after {seconds} s (hung collective or kernel?)"}), flush=True)
        sys.stderr.write(f"bench.py watchdog fired on rank {rank}\n")
        sys.stderr.flush()
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    watchdog = arm_watchdog(rank, int(os.environ.get("B200_BENCH_WATCHDOG_S", "900")))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from transformers import BertConfig
    from spokennlp_b200 import lib, ops
    from spokennlp_b200.trainer import DataParallelTrainer, TopicSegModel

    from spokennlp_b200.blocks import Experimental
    opt_in = [n for n in Experimental.active() if n not in Experimental.DEFAULT.split(",")]      # variants beyond the default set
    off = [n for n in Experimental.DEFAULT.split(",") if n not in Experimental.active()]
    torch.manual_seed(0)
    cfg = BertConfig(hidden_dropout_prob=args.dropout, attention_probs_dropout_prob=args.dropout, **CFG)
    model = TopicSegModel(cfg)
    trainer = DataParallelTrainer(model, lr=5e-5, total_steps=10 * (args.steps + args.warmup) + 1000)
    host = [t.pin_memory() for t in synth_batch(torch, BATCH, SEQ, 1234 + rank)]
    dev = [t.cuda(non_blocking=True) for t in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) * 1e-3

    # The whole step, NCCL collectives included, is replayed from one CUDA graph at every GPU count (B200_DP_GRAPH=0 forces the
    # eager step: the same kernels and collectives launched one by one, 3 % slower on 2 GPUs in round 1).
    dp_graph = os.environ.get("B200_DP_GRAPH", "auto")
    use_graph = not args.no_graph and dp_graph != "0"
    graphed = trainer.capture(*dev) if use_graph else False
    # every timed loop of this file starts the same way: an idle part (set-up and graph capture leave it warm), then the warm-up steps
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    time.sleep(1.0)
    for _ in range(max(3, args.warmup)):
        trainer.step(*dev)
    l0 = lib.launch_count()
    sampler.begin()
    secs = timed(lambda: trainer.step(*dev), args.steps)
    launches = lib.launch_count() - l0
    if graphed:       # replayed launches are not seen by the host-side counter: kernels per captured step x steps
        launches = trainer.kernels_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = world * BATCH * args.steps / secs

    # end to end: pinned host batch -> H2D every step, loss read back every step
    h2d = sum(t.numel() * t.element_size() for t in host)

    e2e_losses = []

    def e2e_step():
        # pinned host batch -> HBM, the step, and the D2H read of this step's loss statistics are all inside the timed
        # region; the host consumes each loss one step late (DataParallelTrainer.step_from_host), so it never idles the GPU
        e2e_losses.append(trainer.step_from_host(*host))
    sampler2 = ClockSampler(local)
    if rank == 0:
        sampler2.start()                       # nvidia-smi needs a few hundred ms before its first sample
    torch.cuda.synchronize()
    time.sleep(1.0)                            # same starting point as the resident-batch loop above: an idle part, then 3 warm-up steps
    for _ in range(3):
        e2e_step()
    trainer.drain()
    del e2e_losses[:]

    def e2e_run():
        for _ in range(args.steps):
            e2e_step()
        e2e_losses.append(trainer.drain())     # the last step's loss is read inside the timed region as well
    sampler2.begin()
    e2e_secs = timed(e2e_run, 1)
    e2e_clocks = sampler2.stop() if rank == 0 else None
    e2e_value = world * BATCH * args.steps / e2e_secs
    assert sum(l is not None for l in e2e_losses) == args.steps, "every timed step's loss must have been read back"
    loss = trainer.loss_value()

    # the kernel table (roofline object): every hot kernel timed ALONE against the measured BURST peaks — taken here, before the
    # sustained run below leaves the part at its power cap for seconds (after it the same kernels time 8-10 % slower)
    kr = None
    if rank == 0:
        peaks, peak_src = load_peaks()
        torch.cuda.synchronize()
        time.sleep(1.5)                        # (the burst peaks were taken on an idle part: let the power state of the timed loops decay)
        kr = kernel_rooflines(torch, ops, lib, peaks, args.dropout)

    # sustained: the same resident-batch step for >= args.sustained_s seconds with its own clock sample, so that the fraction of
    # the SUSTAINED peak is measured under sustained clocks (the 20-step headline lasts 0.3 s and runs at burst clocks)
    sustained = None
    if args.sustained_s > 0:
        n_sus = max(args.steps, int(args.sustained_s / (secs / args.steps)) + 1)
        sampler3 = ClockSampler(local)
        if rank == 0:
            sampler3.start()
        sus_secs = timed(lambda: trainer.step(*dev), n_sus)
        sus_clocks = sampler3.stop() if rank == 0 else None
        sustained = {"value": world * BATCH * n_sus / sus_secs, "unit": "seq/s", "steps": n_sus, "seconds": sus_secs,
                     "ms_per_step": sus_secs / n_sus * 1e3, "clocks": sus_clocks}

    # padded: SURVEY.md §8d's second input variant — right-padded windows with valid lengths ~U[256, 512] (what the reference's
    # collator produces for the last window of a document).  Same step on the same shapes; attention skips fully masked key blocks.
    padded = None
    if args.padded:
        phost = synth_batch(torch, BATCH, SEQ, 4321 + rank, padded=True)
        pdev = [t.cuda() for t in phost]
        fill = float(phost[1].float().mean())
        for _ in range(3):                     # (a captured step replays as is: masks and lengths are read from device memory)
            trainer.step(*pdev)
        p_secs = timed(lambda: trainer.step(*pdev), args.steps)
        # packed: only the valid tokens go through the encoder; the packed row count is read by the host
        # once per step, so these steps are launched eagerly
        for _ in range(3):
            trainer.step(*pdev, pack=True)
        k_secs = timed(lambda: trainer.step(*pdev, pack=True), args.steps)
        padded = {"value": world * BATCH * args.steps / k_secs, "unit": "seq/s", "ms_per_step": k_secs / args.steps * 1e3,
                  "mode": "packed rows (cu_seqlens attention), eager launches",
                  "mean_fill": fill, "tokens_per_s": world * BATCH * SEQ * fill * args.steps / k_secs,
                  "ideal_if_padding_were_free": value / fill, "frac_of_ideal": (world * BATCH * args.steps / k_secs) / (value / fill),
                  "padding_computed": {"value": world * BATCH * args.steps / p_secs, "ms_per_step": p_secs / args.steps * 1e3,
                                       "mode": "padded rows computed, fully masked key blocks skipped, CUDA graph" if graphed else "padded rows computed, eager"},
                  "note": "rows are right-padded windows with valid lengths ~U[256,512]; seq/s counts windows"}

    # the gradient exchange alone (N > 1): one layer bucket and the trailing embeddings bucket, NCCL timed with CUDA events,
    # bus bandwidth = 2 (N-1)/N x bytes / time (the figure nccl-tests reports; 725 GB/s measured on this pool at 1 GiB)
    nccl = None
    if world > 1:
        nccl = {}
        g32 = trainer.flat.grad32
        for name, (lo, hi) in (("layer_bucket", trainer.layer_slices[0]), ("embeddings_bucket", trainer.emb_slice)):
            buf = torch.zeros(hi - lo, dtype=g32.dtype, device=g32.device)
            for grp_name, grp in (("default", None), ("background", trainer.bg_group)):
                if grp_name == "background" and grp is None:
                    continue
                t = timed(lambda: dist.all_reduce(buf, group=grp), 10) / 10
                nccl[f"{name}_{grp_name}"] = {"bytes": buf.numel() * 4, "ms": t * 1e3, "busbw_GBps": 2 * (world - 1) / world * buf.numel() * 4 / t / 1e9}
        nccl["comm_ctas_background"] = trainer.comm_ctas

    if rank == 0:
        dom = kr["gemm_ffn_up_gelu"]
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_kernel.json")))
        except Exception:  # noqa: BLE001
            pass
        cpu = time_cpu_reference(torch, steps=3, warmup=1, batch=2, dropout=args.dropout) if (world == 1 and not args.no_cpu_baseline) else None
        line = {
            "metric": "512-tok seq/sec BERT-base topic-seg fine-tune", "value": value, "unit": "seq/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16 operands, fp32 accumulate/master", "data": "synthetic",
            "config": {"workload": "emnlp2023-topic_segmentation BERT-base fine-tune, 512-tok windows, bsz 32/GPU "
                                   "(fwd + bwd + grad allreduce + clip + AdamW)", "seq_len": SEQ, "batch_per_gpu": BATCH,
                       "global_batch": BATCH * world, "parallelism": f"dp{world}", "dropout": args.dropout, "cuda_graph": bool(graphed),
                       "opt_in_variants": opt_in,        # B200_EXP; [] = the default, GPU-validated kernels
                       "disabled_default_variants": off,
                       "l2": "per-step working set ~5.4 GB of activations >> 126 MB L2 (no explicit flush needed)"},
            "encoder_flop_util": {"flop_per_seq": FLOP_PER_SEQ, "achieved_tflops_per_gpu": value / world * FLOP_PER_SEQ / 1e12,
                                  "peak_tflops_sustained": peaks["bf16_tflops_sustained"], "peak_source": peak_src,
                                  "frac_of_sustained": value / world * FLOP_PER_SEQ / 1e12 / peaks["bf16_tflops_sustained"]},
            # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this kernel per launch, READ from the committed artefact of
            # an `ncu --set full` capture (profiles/roofline_kernel.json, written by tools/roofline_traffic.py); null if absent.
            # Algorithmic bytes: 25.2 MB x + 4.7 MB W read, 2 x 100.7 MB written (gelu(z) and gelu'(z): both are training outputs).
            "roofline": {"kernel": "gemm2_f16_kernel<256,K-major,K-major,BIAS_GELU> (FFN-up 16384x3072x768, 2-CTA tcgen05)",
                         "bound": "tensor", "achieved": dom["achieved"], "peak": dom["peak"], "unit": "TFLOP/s", "frac": dom["frac"],
                         "traffic": traffic.get("dram_bytes"), "traffic_source": traffic.get("source"),
                         "algorithmic_bytes": 2 * BATCH * SEQ * 768 + 2 * 768 * 3072 + 2 * 2 * BATCH * SEQ * 3072,
                         "peak_source": peak_src + " bf16 burst (kernel timed alone)"},
            "kernels": kr,
            "e2e": {"value": e2e_value, "unit": "seq/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "ms_per_step": e2e_secs / args.steps * 1e3, "sm_mhz": (e2e_clocks or {}).get("sm_mhz")},
            "gpu_launches": launches, "clocks": clocks, "final_loss": loss,
        }
        if sustained is not None:
            sustained["frac_of_sustained_peak"] = sustained["value"] / world * FLOP_PER_SEQ / 1e12 / peaks["bf16_tflops_sustained"]
            line["sustained"] = sustained
        if padded is not None:
            line["padded"] = padded
        if nccl is not None:
            line["nccl"] = nccl
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    watchdog.cancel()
    if world > 1:
        # Tear-down: a captured step holds NCCL work; destroying the communicator (or letting the interpreter run the
        # destructors) while the graph is alive can block for ever.  Drop the graph, drain the device, meet the peers once
        # more, then leave without running NCCL's tear-down (the process is ending anyway).
        sys.stdout.flush()
        sys.stderr.flush()
        bye = threading.Timer(30.0, lambda: os._exit(0))     # the result is out: never let the farewell barrier hang the job
        bye.daemon = True
        bye.start()
        trainer.release_graph()
        barrier()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1, help="hidden / attention-probability / classifier dropout (reference default 0.1)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph of the step")
    ap.add_argument("--sustained-s", type=float, default=3.0, help="length of the extra sustained-clock run in seconds (0 = skip)")
    ap.add_argument("--no-padded", dest="padded", action="store_false", help="skip the padded-window sub-record")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torch.distributed.run, one rank per GPU (what the driver does itself)
        import socket
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
