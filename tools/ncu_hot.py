"""Hot instructions / stall reasons of one kernel from an .ncu-rep (needs -lineinfo builds for SASS->source):
    python tools/ncu_hot.py gpurun_out/prof.ncu-rep attn_fwd_kernel [launch_index]
"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# split per launch: each block starts with a "Kernel Name" row
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
    elif cur is not None:
        cur.append(r)
blk = blocks[which]
hdr, data = blk[0], [r for r in blk[1:] if len(r) == len(blk[0])]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
print(f"{kern} launch {which}: {len(blocks)} launches in report, {tot} samples, {len(data)} SASS instructions")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    s = int(r[ix["# Samples"]])
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{s:6d} {100 * s / max(tot, 1):5.1f}% exec={r[ix['Instructions Executed']]:>8} {r[ix['Source']].strip()[:80]:80s} {st}")
