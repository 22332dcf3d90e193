#!/usr/bin/env python
"""Turn the files the prepared round-2 GPU calls leave in gpurun_out/ (tools/round2_call1.sh, round2_scaling.sh) into one
markdown summary:  python tools/summarize_round2.py [gpurun_out] > profiles/r02_first_calls.md"""
import glob
import json
import os
import sys

d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"


def last_json(path):
    try:
        lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else None
    except Exception:
        return None


def tail(path, n=3):
    try:
        return [l.rstrip() for l in open(path, errors="replace").read().strip().splitlines()[-n:]]
    except Exception:
        return []


print("# Round 2, prepared GPU calls — summary\n")
for name in ("r2_gpu_suite.log", "r2_experimental.log", "r2_experimental_isolated.log", "r2_sanitizer_memcheck.log"):
    t = tail(os.path.join(d, name), 2)
    if t:
        print(f"* `{name}`: " + " / ".join(x.strip() for x in t if x.strip()))
print()

rows = []
for p in sorted(glob.glob(os.path.join(d, "r2_bench_*.json"))):
    j = last_json(p)
    tag = os.path.basename(p)[len("r2_bench_"):-5]
    if j and j.get("value"):
        k = j.get("kernels", {})
        rows.append((tag, j["value"], j["ms_per_step"], j["e2e"]["value"], j.get("final_loss"), (j.get("clocks") or {}).get("sm_mhz"),
                     k.get("gemm_ffn_down_res", {}).get("achieved"), k.get("attn_fwd", {}).get("ms"), k.get("attn_bwd", {}).get("ms")))
    else:
        rows.append((tag, None, None, None, None, None, None, None, None))
if rows:
    base = next((r[1] for r in rows if r[0] == "default" and r[1]), None)
    print("## bench.py, one B200, same box (B200_EXP = variant set)\n")
    print("| variants | seq/s | vs default | ms/step | e2e seq/s | final loss | SM MHz | FFN-down TFLOP/s | attn fwd ms | attn bwd ms |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    f = lambda v, nd=1: "-" if v is None else f"{v:.{nd}f}"
    for r in rows:
        rel = "-" if not (base and r[1]) else f"{r[1] / base:.3f}"
        print(f"| `{r[0]}` | {f(r[1])} | {rel} | {f(r[2], 3)} | {f(r[3])} | {f(r[4], 4)} | {f(r[5], 0)} | {f(r[6], 0)} | {f(r[7], 4)} | {f(r[8], 4)} |")
    print()

ab = os.path.join(d, "r2_variants_ab.jsonl")
if os.path.exists(ab):
    print("## kernel A/B (tools/variants_ab.py)\n")
    for line in open(ab):
        try:
            j = json.loads(line)
        except Exception:
            continue
        print("* " + ", ".join(f"{k}={v:.1f}" if isinstance(v, float) else f"{k}={v}" for k, v in j.items()))
    print()

scale = []
for p in sorted(glob.glob(os.path.join(d, "r2_scale_*.json"))):
    j = last_json(p)
    tag = os.path.basename(p)[len("r2_scale_"):-5]
    scale.append((tag, j))
if scale:
    one = next((r[1] for r in rows if r[0] == "default" and r[1]), None)
    print("## data-parallel scaling (tools/round2_scaling.sh)\n")
    print("| run | GPUs | seq/s | ms/step | e2e seq/s | vs N x 1-GPU | graph | note |")
    print("|---|---:|---:|---:|---:|---:|---|---|")
    for tag, j in scale:
        if j and j.get("value"):
            n = j["n_gpus"]
            eff = "-" if not one else f"{j['value'] / (n * one):.3f}"
            print(f"| `{tag}` | {n} | {j['value']:.1f} | {j['ms_per_step']:.3f} | {j['e2e']['value']:.1f} | {eff} | {j['config'].get('cuda_graph')} | |")
        else:
            note = (j or {}).get("error", "no JSON line (timeout?)")
            print(f"| `{tag}` | | | | | | | {note} |")
