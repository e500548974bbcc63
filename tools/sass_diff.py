#!/usr/bin/env python
"""Which kernels differ between two builds of liblgarb.so?  Per-function comparison of `cuobjdump -sass` listings (white
space and fatbin headers ignored).  Used to show that adding an opt-in kernel leaves every MEASURED kernel byte for byte as
it was (adding a caller inside lgar_kernels.cu does not: the device functions are inlined per module by cost).

    python tools/sass_diff.py old.so new.so        # exit 1 if a kernel present in both differs
"""
import re
import subprocess
import sys

SKIP = ("Fatbin", "=====", "arch =", "code version", "host =", "compile_size", "identifier", "code for", ".target")


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], check=True, capture_output=True, text=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
        elif cur:
            t = re.sub(r"\s+", " ", line).strip()
            if t and not t.startswith(SKIP):
                d[cur].append(t)
    return d


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    changed = [k for k in b if k in a and a[k] != b[k]]
    print(f"{len(a)} kernels in {sys.argv[1]}, {len(b)} in {sys.argv[2]}")
    for k in b:
        if k not in a:
            print("  new    ", k)
    for k in a:
        if k not in b:
            print("  gone   ", k)
    for k in changed:
        print("  CHANGED", k, f"({len(a[k])} -> {len(b[k])} lines)")
    if not changed:
        print("  every kernel present in both is identical")
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
