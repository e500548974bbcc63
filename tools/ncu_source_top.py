#!/usr/bin/env python
"""Top source lines of one kernel of an .ncu-rep by warp-stall samples and executed instructions.
usage: ncu_source_top.py report.ncu-rep LAUNCH_INDEX [N]"""
import csv, subprocess, sys, collections, os
rep, idx = sys.argv[1], int(sys.argv[2])
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if len(r) < 10 or r[0] in ("Line No", ""):
        continue
    try:
        ln = int(r[0]); smp = int(r[6] or 0); wi = int(r[7] or 0); ti = int(r[8] or 0)
    except ValueError:
        continue
    a = agg[(cur, ln)]
    a[0] += smp; a[1] += wi; a[2] += ti; a[3] = r[1].strip()[:90]
ts = sum(a[0] for a in agg.values()); tw = sum(a[1] for a in agg.values()); tt = sum(a[2] for a in agg.values())
print(f"total samples {ts}  warp-instr {tw/1e6:.1f}M  lanes/instr {tt/max(tw,1):.1f}")
byfile = collections.Counter()
for (f, l), a in agg.items(): byfile[f] += a[0]
print("samples by file:", dict(byfile))
for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:N]:
    print(f"{100*a[0]/ts:5.1f}% smp {100*a[1]/tw:5.1f}% ins {a[2]/max(a[1],1):5.1f} ln  {f}:{l}  {a[3]}")
