"""Group the source-page samples of one kernel by the instruction's execution count (= which warp role runs it: a role's
per-block instructions all share one count) and by stall reason:
    python tools/ncu_roles.py gpurun_out/x.ncu-rep attn_fwd3_kernel [launch]"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
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
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by = collections.defaultdict(lambda: collections.Counter())
ninst = collections.Counter()
for r in data:
    e = int(r[ix["Instructions Executed"]])
    ninst[e] += 1
    for h in stalls:
        by[e][h[6:]] += int(r[ix[h]])
tot = sum(sum(c.values()) for c in by.values())
print(f"{kern}: {tot} samples")
for e, c in sorted(by.items(), key=lambda kv: -sum(kv[1].values()))[:14]:
    s = sum(c.values())
    print(f"exec={e:8d} ninstr={ninst[e]:5d} samples={s:6d} {100*s/tot:5.1f}%  " + ", ".join(f"{k}={v}" for k, v in c.most_common(7)))
# opcode histogram of the dominant class
top = max(by.items(), key=lambda kv: sum(kv[1].values()))[0]
opmix = collections.Counter()
for r in data:
    if int(r[ix["Instructions Executed"]]) == top:
        src = r[ix["Source"]].strip().split()
        op = src[1] if src and src[0].startswith("@") else (src[0] if src else "?")
        opmix[op.split(".")[0]] += 1
print(f"opcode mix of exec={top} ({ninst[top]} instructions per thread per block):", ", ".join(f"{k}={v}" for k, v in opmix.most_common(25)))
