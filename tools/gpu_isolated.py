"""Run every `-m gpu` test node in its own process with a timeout, so that one trapped / hung kernel (which poisons
the CUDA context of its process) cannot hide the results of the others.  Writes gpurun_out/isolated.log.

    python tools/gpu_isolated.py [pytest -k expression] [--timeout 180]
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
os.makedirs("gpurun_out", exist_ok=True)
args = sys.argv[1:]
timeout = 180
if "--timeout" in args:
    i = args.index("--timeout")
    timeout = int(args[i + 1])
    del args[i:i + 2]
files = [a for a in args if a.endswith(".py")] or ["tests"]
kexpr = [a for a in args if not a.endswith(".py")]
cmd = [sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu"] + files
if kexpr:
    cmd += ["-k", kexpr[0]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
nodes = [l.strip() for l in out.splitlines() if "::" in l]
log = open("gpurun_out/isolated.log", "a")
log.write(f"==== {time.strftime('%H:%M:%S')} {len(nodes)} nodes\n")
npass = 0
for n in nodes:
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", n, "--no-header", "-p", "no:cacheprovider"],
                           capture_output=True, text=True, timeout=timeout)
        ok = r.returncode == 0
        tail = (r.stdout + r.stderr)
    except subprocess.TimeoutExpired as e:
        ok = False
        tail = "TIMEOUT\n" + ((e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or ""))
    npass += ok
    line = f"{'PASS' if ok else 'FAIL'} {time.time() - t0:6.1f}s {n}"
    print(line, flush=True)
    log.write(line + "\n")
    if not ok:
        keep = [l for l in tail.splitlines() if l.strip()][-60:]
        log.write("\n".join("    " + l for l in keep) + "\n")
        print("\n".join("    " + l for l in keep[-25:]), flush=True)
    log.flush()
print(f"{npass}/{len(nodes)} passed")
log.write(f"{npass}/{len(nodes)} passed\n")
