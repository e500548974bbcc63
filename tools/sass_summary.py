#!/usr/bin/env python
"""Instruction mix of every kernel of liblgarb.so from `cuobjdump -sass` (static counts: what the code consists of, not what a
run executes).  Columns: SASS instructions, bytes of code (16 per instruction), FP64 arithmetic (DFMA/DMUL/DADD/DSETP/MUFU.RCP64H
...), global loads/stores by width (LDG/STG .64/.128/.ENL2.256), local loads/stores (LDL/STL: the thread-private lane record and
spills), shared (LDS/STS), atomics, warp votes/shuffles/match, branches/calls.  Device functions that are not inlined
(`__noinline__`) are separate symbols in the listing and are attributed to the kernel group that calls them by name only in the
text (they are listed as functions of the module).

    python tools/sass_summary.py [lib.so] > profiles/r2_sass_summary.md"""
import collections
import hashlib
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "lgar_c_b200" / "liblgarb.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], check=True, capture_output=True, text=True).stdout
funcs = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        funcs[cur][m.group(1)] += 1


def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], check=True, capture_output=True, text=True).stdout.strip()
    except Exception:
        return n


def cls(c):
    g = collections.Counter()
    for op, k in c.items():
        base = op.split(".")[0]
        g["total"] += k
        if base in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX") or op.startswith("MUFU.RCP64") or op.startswith("MUFU.RSQ64"):
            g["fp64"] += k
        elif base in ("LDG", "LD"):
            g["ldg256" if "256" in op else "ldg128" if "128" in op else "ldg"] += k
        elif base in ("STG", "ST"):
            g["stg256" if "256" in op else "stg128" if "128" in op else "stg"] += k
        elif base == "LDL":
            g["ldl"] += k
        elif base == "STL":
            g["stl"] += k
        elif base in ("LDS", "STS", "LDSM"):
            g["shared"] += k
        elif base in ("ATOM", "ATOMG", "RED", "ATOMS"):
            g["atomic"] += k
        elif base in ("VOTE", "VOTEU", "SHFL", "MATCH", "WARPSYNC", "BSSY", "BSYNC"):
            g["warp"] += k
        elif base in ("BRA", "BRX", "CALL", "RET", "JMP", "EXIT"):
            g["branch"] += k
        elif base in ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS"):
            g["bulk_async"] += k
    return g


print("# Round 2: SASS instruction mix per kernel (`python tools/sass_summary.py`)\n")
print(f"`{lib.name}` sha256 `{hashlib.sha256(lib.read_bytes()).hexdigest()[:16]}...`, `cuobjdump -sass`, static instruction counts.\n")
print("| function | instr | code KB | FP64 | LDG (.128 / .256) | STG (.128 / .256) | LDL | STL | shared | atomics | vote/shfl/sync | branch/call |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
rows = []
for n, c in funcs.items():
    g = cls(c)
    d = demangle(n)
    d = re.sub(r"\((?!int\)|bool\)).*$", "", d).replace("void ", "").replace("(anonymous namespace)::", "")
    d = d.replace("(int)", "").replace("(bool)0", "plain").replace("(bool)1", "fast-path")
    for k, v in (("<0,", "<FETCH,"), ("<1,", "<HEAD,"), ("<2,", "<FRONT,"), ("<3,", "<SEARCH,"), ("<4,", "<TAIL,"), ("<5,", "<CORR,")):
        d = d.replace("lgar_wave_kernel" + k, "lgar_wave_kernel" + v)
    rows.append((g["total"], d, g))
for tot, d, g in sorted(rows, key=lambda r: -r[0]):
    print(f"| `{d}` | {tot} | {tot * 16 / 1024:.1f} | {g['fp64']} | {g['ldg']} ({g['ldg128']} / {g['ldg256']}) | {g['stg']} ({g['stg128']} / {g['stg256']}) | "
          f"{g['ldl']} | {g['stl']} | {g['shared']} | {g['atomic']} | {g['warp']} | {g['branch']} |")
print("\nNo `UBLKCP` / `UTMALDG` / `LDGSTS` (bulk asynchronous copies) anywhere: the lane records move with per-lane 256-bit `LDG.E.ENL2.256` / `STG.E.ENL2.256`"
      " (one full 32-byte sector per lane per access) into thread-private (`STL`/`LDL`) copies, or -- finishing kernel -- into shared memory.")
