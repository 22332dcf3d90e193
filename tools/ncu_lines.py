"""Per-source-line stall samples of one kernel from an .ncu-rep (built with -lineinfo, captured with --import-source on):
    python tools/ncu_lines.py rep.ncu-rep kernel_regex [launch_index] [top_n]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# sections: each starts with a "File Path" row, then "Function Name", then header "Line No",...
secs, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
launch = -1
seen_kernel_files = {}
out = []
for s in secs:
    hdr = next((r for r in s["rows"] if r and r[0] == "Line No"), None)
    if hdr is None:
        continue
    ix = {}
    for i, h in enumerate(hdr):
        ix.setdefault(h, i)
    fn = next((r[1] for r in s["rows"] if r and r[0] == "Function Name"), "")
    key = (fn, s["file"])
    seen_kernel_files[key] = seen_kernel_files.get(key, -1) + 1
    if seen_kernel_files[key] != which:
        continue
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in s["rows"]:
        if len(r) != len(hdr) or not r[0].isdigit():
            continue
        n = int(r[ix["# Samples"]])
        if n == 0:
            continue
        st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:3]
        out.append((n, s["file"].split("/")[-1], int(r[0]), r[1].strip()[:90], int(r[ix["Instructions Executed"]]), st))
tot = sum(o[0] for o in out)
print(f"{kern}: {tot} samples attributed to source lines")
for n, f, ln, src, ex, st in sorted(out, reverse=True)[:top]:
    print(f"{n:6d} {100 * n / tot:5.1f}% {f}:{ln:<4d} exec={ex:>9d} {src:90s} {[(a, b) for a, b in st if a]}")
