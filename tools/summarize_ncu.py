"""Turn ncu artefacts brought back in gpurun_out/ into small tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launch_list.md
    python tools/summarize_ncu.py full gpurun_out/prof_r01.ncu-rep profiles/r01_kernel_metrics.md
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot, n = 0.0, 0
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        t = t / 1e3 if unit == "ns" else (t * 1e3 if unit == "ms" else t)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
        n += 1
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over {n} consecutive launches of "
                f"`bench.py` (about one fine-tuning step; cold-cache, serialised: compare SHARES). Total {tot / 1e3:.2f} ms.\n\n")
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k.strip()[:100]}` | {c} | {t:.1f} | {100 * t / tot:.1f}% | {t / c:.1f} |\n")
    print("wrote", dst)


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\nOne row per captured launch (bench shapes: B=32, S=512, BERT-base).\n\n")
        f.write("| kernel | " + " | ".join(lbl for _, lbl in METRICS) + " |\n|---|" + "---:|" * len(METRICS) + "\n")
        for d in data:
            name = re.sub(r"\(.*", "", d[idx["Kernel Name"]]).strip()
            if name.startswith("void at::"):
                continue
            cells = []
            for m, _ in METRICS:
                if m in idx:
                    v = d[idx[m]]
                    try:
                        v = f"{float(v.replace(',', '')):.1f}"
                    except Exception:
                        pass
                    cells.append(f"{v} {units[idx[m]]}".strip())
                else:
                    cells.append("-")
            f.write(f"| `{name[:70]}` | " + " | ".join(cells) + " |\n")
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
