"""The HBM-bound kernels of the path at their BASELINE shapes, each launched twice, for one `ncu --set full` capture:
LayerNorm fwd / bwd, embeddings fwd / bwd, cls head fwd / bwd (B=32, S=512, H=768), PoNet pooling mixer fwd / bwd ([2, 4096], 12 heads).

    ncu --set full --clock-control none -k regex:"ln_|embed_ln|cls_head|ponet_" -f -o gpurun_out/hbm python tools/prof_hbm.py
    python tools/prof_hbm.py summarize gpurun_out/hbm.ncu-rep profiles/r02_hbm_rooflines.md      (here, no GPU)
The summary compares the measured duration with SURVEY.md §8d's ALGORITHMIC bytes for each kernel."""
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B, S, H, heads, V = 32, 512, 768, 12, 30523
M = B * S
PB, PS = 2, 4096
# algorithmic bytes per launch (fp16 activations, fp32 residual stream / tables)
ALGO = {
    "ln_fwd_kernel": ("read fp32 pre-LN sum, write fp16 + fp32 copies", M * H * (4 + 2 + 4)),
    "ln_bwd2_kernel": ("read dy fp16 + x fp32, write dx fp16 (+ masked copy under dropout)", M * H * (2 + 4 + 2 + 2)),
    "embed_ln_fwd_kernel": ("gather fp32 word rows + positions, write fp16 + fp32", M * H * (4 + 2 + 4) + S * H * 4),
    "embed_ln_bwd_kernel": ("read dy fp16, re-gather word rows, scatter-add fp32 word + position gradients", M * H * (2 + 4 + 4 + 4)),
    "cls_head_fwd_kernel": ("read fp16 activations, write 2 fp32 logits per row", M * H * 2 + M * 8),
    "cls_head_bwd_kernel": ("read fp16 activations, write fp16 dh", M * H * (2 + 2)),
    "ponet_mix_kernel": ("SURVEY §8d: read K, Q, O, Sg, Lc once + write out = 6 S H 2 per sequence (all mixer kernels together)", PB * 6 * PS * H * 2),
    "ponet_bwd_rows_kernel": ("read proj (5 S H 2) + dout, write dproj (5 S H 2) per sequence (all backward kernels together)", PB * 11 * PS * H * 2),
}


def run():
    import torch
    from spokennlp_b200 import ops
    dev, f16 = "cuda", torch.float16
    torch.manual_seed(0)
    seed = torch.tensor([7], dtype=torch.int32, device=dev)
    drop = ops.Dropout(seed, 2, 0.1)
    pre = torch.randn(M, H, device=dev)
    g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    y, y32 = torch.empty(M, H, device=dev, dtype=f16), torch.empty(M, H, device=dev)
    mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
    dx, dxd = torch.empty(M, H, device=dev, dtype=f16), torch.empty(M, H, device=dev, dtype=f16)
    dg, db, dbias = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
    ids = torch.randint(1000, V - 1, (M,), device=dev)
    word, pos_tab, type_tab = torch.randn(V, H, device=dev) * 0.02, torch.randn(512, H, device=dev) * 0.02, torch.randn(2, H, device=dev) * 0.02
    dword, dpos, dtype_tab = torch.zeros_like(word), torch.zeros_like(pos_tab), torch.zeros_like(type_tab)
    W, bc = torch.randn(2, H, device=dev) * 0.02, torch.zeros(2, device=dev)
    labels = torch.where(torch.rand(M, device=dev) < 0.05, torch.randint(0, 2, (M,), device=dev), torch.full((M,), -100, device=dev))
    stats = torch.zeros(2, device=dev)
    dW, dbc = torch.zeros_like(W), torch.zeros(2, device=dev)
    one = torch.ones(1, device=dev)
    # PoNet mixer
    PH = heads * 64
    proj = torch.randn(PB * PS, 5 * PH, device=dev, dtype=f16)
    seg = (torch.arange(PS, device=dev) // 24 + 1).repeat(PB, 1).contiguous()
    mix = torch.empty(PB * PS, PH, device=dev, dtype=f16)
    dmix = torch.randn(PB * PS, PH, device=dev, dtype=f16)
    dproj = torch.empty_like(proj)
    for _ in range(2):
        ops.layernorm_fwd(pre, g, b, 1e-12, y=y, y32=y32, mean=mean, rstd=rstd)
        ops.layernorm_bwd(y, pre, mean, rstd, g, dx, dg, db, dbias=dbias, dx_drop=dxd, drop=drop)
        ops.embed_ln_fwd(ids, None, None, None, word, pos_tab, type_tab, g, b, 1e-12, M, S, H, y=y, y32=y32)
        ops.embed_ln_bwd(y, None, ids, None, None, word, pos_tab, type_tab, g, dword, dpos, dtype_tab, dg, db, None, 1e-12, M, S, H, pad_id=0)
        logits = ops.cls_head_fwd(y, W, bc)
        stats.zero_()
        ops.ce_stats(logits, labels, stats)
        ops.cls_head_bwd(y, logits, labels, stats, W, dx, dW, dbc, scale=one)
        ws = ops.ponet_mix_fwd(proj, seg, mix, PB, PS, heads, PS + 2)
        ops.ponet_mix_bwd(proj, dmix, seg, ws, dproj, PB, PS, heads, PS + 2)
    torch.cuda.synchronize()
    print("done")


def summarize(rep, dst):
    import json
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    def val(d, name, scale=True):
        v, u = float(d[ix[name]].replace(",", "")), units[ix[name]].lower()
        return v * ({"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1) if scale else 1)
    per = {}
    for d in data:                              # keep the LAST launch of every kernel (second iteration)
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", d[ix["Kernel Name"]])).replace("void ", "").replace("b200::", "").strip()
        per.setdefault(name, []).append(d)
    # PoNet: the mixer is several kernels; report them together
    groups = {"ponet_mix (fwd, all kernels)": [k for k in per if k.startswith("ponet_") and "bwd" not in k and "lse" not in k],
              "ponet_mix (bwd, all kernels)": [k for k in per if k.startswith("ponet_") and ("bwd" in k or "lse" in k)]}
    with open(dst, "w") as f:
        f.write(f"# HBM-bound kernels under `ncu --set full` ({os.path.basename(rep)})\n\nShapes: B=32, S=512, H=768 (BERT-base bench shape); PoNet mixer [2, 4096], 12 heads.  "
                f"`achieved` = ALGORITHMIC bytes (SURVEY.md §8d) / kernel duration under ncu (cold cache, serialised); peak = {peak:.0f} GB/s (MEASURED_PEAKS.json).  "
                "`DRAM` = dram__bytes_read.sum + dram__bytes_write.sum of the launch.\n\n")
        f.write("| kernel | duration us | algorithmic MB | achieved GB/s | frac of HBM peak | DRAM MB | DRAM throughput % | what the bytes are |\n|---|---:|---:|---:|---:|---:|---:|---|\n")
        for name, ds in per.items():
            if name.startswith("ponet_") or name not in ALGO:
                continue
            d = ds[-1]
            t = val(d, "gpu__time_duration.sum")
            what, nbytes = ALGO[name]
            dram = val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum")
            f.write(f"| `{name}` | {t:.1f} | {nbytes / 1e6:.1f} | {nbytes / t / 1e3:.0f} | {nbytes / t / 1e3 / peak:.2f} | {dram / 1e6:.1f} | "
                    f"{val(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', False):.0f} | {what} |\n")
        for gname, ks in groups.items():
            if not ks:
                continue
            t = sum(val(per[k][-1], "gpu__time_duration.sum") for k in ks)
            dram = sum(val(per[k][-1], "dram__bytes_read.sum") + val(per[k][-1], "dram__bytes_write.sum") for k in ks)
            what, nbytes = ALGO["ponet_mix_kernel" if "fwd" in gname else "ponet_bwd_rows_kernel"]
            f.write(f"| `{gname}`: " + ", ".join(f"{k} {val(per[k][-1], 'gpu__time_duration.sum'):.1f}" for k in ks) +
                    f" | {t:.1f} | {nbytes / 1e6:.1f} | {nbytes / t / 1e3:.0f} | {nbytes / t / 1e3 / peak:.2f} | {dram / 1e6:.1f} | - | {what} |\n")
    print(open(dst).read())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "summarize":
        summarize(sys.argv[2], sys.argv[3])
    else:
        run()
