"""Write profiles/roofline_kernel.json — the DRAM traffic of bench.py's `roofline` kernel per launch — from an `ncu --set full`
capture, so that the bench line READS the number instead of carrying a literal:

    ncu --set full --clock-control none -k regex:gemm2_f16_kernel -c 4 -f -o gpurun_out/gelu python tools/roofline_traffic.py run
    python tools/roofline_traffic.py summarize gpurun_out/gelu.ncu-rep            (here, no GPU needed)
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import torch
    from spokennlp_b200 import ops
    M, H, I = 32 * 512, 768, 3072
    x = torch.randn(M, H, device="cuda", dtype=torch.float16)
    w1 = torch.randn(I, H, device="cuda", dtype=torch.float16) * 0.02
    b1 = torch.zeros(I, device="cuda")
    h, z = torch.empty(M, I, device="cuda", dtype=torch.float16), torch.empty(M, I, device="cuda", dtype=torch.float16)
    for _ in range(4):
        ops.gemm(x, w1, h, epilogue=ops.EPI_BIAS_GELU, bias=b1, out2=z)
    torch.cuda.synchronize()
    print("done")


def summarize(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(d, name):
        v, u = float(d[ix[name]].replace(",", "")), units[ix[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    sel = [d for d in data if "gemm2_f16_kernel" in d[ix["Kernel Name"]] and ", 2, __half" in d[ix["Kernel Name"]].replace("(int)", "")] or \
          [d for d in data if "gemm2_f16_kernel" in d[ix["Kernel Name"]]]
    d = sel[-1]                                   # the last captured launch (warm instruction cache, cold data: operands >> L2)
    rd, wr = val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum")
    out = {"kernel": d[ix["Kernel Name"]][:120], "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
           "duration_us_under_ncu": float(d[ix["gpu__time_duration.sum"]].replace(",", "")) / (1e3 if units[ix["gpu__time_duration.sum"]] == "ns" else 1),
           "tensor_pipe_active_pct": float(d[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
           "source": f"ncu --set full, {os.path.basename(rep)} (tools/roofline_traffic.py)"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_kernel.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    run() if sys.argv[1] == "run" else summarize(sys.argv[2])
