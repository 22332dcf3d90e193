#!/usr/bin/env python
"""Compare the SASS of two builds of libb200enc.so kernel by kernel (CPU-only check that a source change left the
machine code of already-validated kernels untouched).

    cuobjdump -sass old.so > old.sass ; cuobjdump -sass new.so > new.sass
    python tools/sass_diff.py old.sass new.sass
"""
import re
import sys


def functions(path):
    out, name, body = {}, None, []
    for line in open(path, errors="replace"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = body
            name, body = m.group(1), []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            body.append(re.sub(r"\s+", " ", line.strip()))
        elif name and re.match(r"\s+/\* 0x[0-9a-f]{16} \*/", line):
            body.append(line.strip())
    if name:
        out[name] = body
    # anonymous-namespace symbols carry a per-build hash
    return {re.sub(r"_GLOBAL__N__[0-9a-f]+_", "_GLOBAL__N__", k): v for k, v in out.items()}


def main():
    a, b = functions(sys.argv[1]), functions(sys.argv[2])
    same = changed = 0
    # a kernel whose template gained a defaulted parameter keeps its code but changes its mangled name
    renamed = {}
    for k in sorted(set(a) - set(b)):
        for k2 in sorted(set(b) - set(a)):
            if k2 not in renamed.values() and a[k] == b[k2] and k2.startswith(k[:k.index("I")]):
                renamed[k] = k2
                break
    for k in sorted(set(a) | set(b)):
        if k in renamed:
            same += 1
            print("RENAMED  ", k, "->", renamed[k], "(identical code)")
        elif k in renamed.values():
            continue
        elif k not in a:
            print("NEW      ", k, len(b[k]))
        elif k not in b:
            print("REMOVED  ", k)
        elif a[k] != b[k]:
            changed += 1
            print("CHANGED  ", k, len(a[k]), "->", len(b[k]))
        else:
            same += 1
    print(f"{same} identical, {changed} changed")
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
