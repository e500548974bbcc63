#!/usr/bin/env python
"""profiles/r2_window_profile.json from the ncu launch list of one timed window (tools/prof_window.sh): measured DRAM bytes per
column-timestep and the share of every kernel in the window's device time, keyed to the sha256 of the liblgarb.so the capture
was taken with (bench.py prints them only when the running library is that build).
usage: make_window_profile.py window_launches.csv COLUMNS HOURS [lib.so] > profiles/r2_window_profile.json"""
import collections
import csv
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
src, ncol, hours = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lib = Path(sys.argv[4]) if len(sys.argv) > 4 else ROOT / "lgar_c_b200" / "liblgarb.so"
lines = [l for l in open(src) if not l.startswith("==")]
unit = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = {}
for row in csv.DictReader(lines):
    per.setdefault(int(row["ID"]), {"name": row["Kernel Name"]})[row["Metric Name"]] = float(row["Metric Value"].replace(",", "")) * unit.get(row["Metric Unit"], 1.0)
stage = {"0": "FETCH", "1": "HEAD", "2": "FRONT", "3": "SEARCH", "4": "TAIL", "5": "CORR"}
by = collections.OrderedDict()
for d in per.values():
    n = d["name"].split("(")[0].replace("void ", "").strip()
    if "lgar_wave_kernel<" in n:
        n = "lgar_wave_kernel<%s> (%s)" % (n.split("<")[1].split(">")[0].replace("(int)", ""), stage.get(n.split("<")[1].replace("(int)", "")[0], "?"))
    a = by.setdefault(n, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
T = sum(a[1] for a in by.values())
B = sum(a[2] + a[3] for a in by.values())
units = ncol * hours
print(json.dumps({
    "lib_sha256": hashlib.sha256(lib.read_bytes()).hexdigest(),
    "source_sha256": __import__("bench").kernel_source_sha256(),
    "source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over every launch of one timed window "
              f"({ncol} columns x {hours} h; tools/prof_window.sh; launch list profiles/r2_window_launches.csv)",
    "columns": ncol, "hours": hours, "launches": sum(a[0] for a in by.values()), "window_ms_under_ncu": T * 1e3,
    "dram_bytes_per_window": B, "dram_bytes_per_column_timestep": B / units,
    "stage_time_share": {n: a[1] / T for n, a in by.items()},
    "per_kernel": {n: {"launches": a[0], "ms": a[1] * 1e3, "dram_read_gb": a[2] / 1e9, "dram_write_gb": a[3] / 1e9,
                       "dram_bytes_per_column_timestep": (a[2] + a[3]) / units} for n, a in by.items()},
}, indent=1))
