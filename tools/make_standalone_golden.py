#!/usr/bin/env python
"""Golden output of the reference's OWN standalone executable (oracle/_ref/lasam_standalone, the unmodified
src/bmi_main_lgar.cxx built by oracle/Makefile) for the test of lgar_c_b200/lasam_b200_standalone: the first
HEAD forcing steps of data_variables.csv and data_layers.csv of the Phillipsburg_with_reservoir configuration.

Run in the build container:  python tools/make_standalone_golden.py"""
import subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.oracle_py import DATA, absolutize_config

HEAD = 400
exe = ROOT / "oracle" / "_ref" / "lasam_standalone"
cfg = absolutize_config(DATA / "configs" / "config_lasam_Phillipsburg_with_reservoir.txt")
with tempfile.TemporaryDirectory() as d:
    subprocess.run([str(exe), str(cfg)], cwd=d, check=True, stdout=subprocess.DEVNULL)
    v = (Path(d) / "data_variables.csv").read_text().splitlines()
    l = (Path(d) / "data_layers.csv").read_text().splitlines()
out = ROOT / "tests" / "golden"
(out / "standalone_Phillipsburg_with_reservoir_variables_head.csv").write_text("\n".join(v[:HEAD + 1]) + "\n")
(out / "standalone_Phillipsburg_with_reservoir_layers_head.csv").write_text("\n".join(l[:2 * HEAD]) + "\n")
print("rows", len(v) - 1, "kept", HEAD)
