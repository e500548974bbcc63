#!/usr/bin/env python
"""Copy the reference's *data* files (soil tables, forcing series, example configs -- no source
code) from /root/reference into data/, so that tests and bench.py can run on the GPU box where
/root/reference does not exist.  Files are byte-identical copies: the soil-table parser quirk
(names containing a blank) must see the original bytes.

    python tools/fetch_reference_data.py [/root/reference]
"""
import shutil
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent.parent / "data"

COPY = {
    "": ["data/vG_default_params.dat", "data/vG_params_stat_nom_ordered.dat",
         "data/vG_default_params_synth_1.dat", "data/vG_default_params_synth_2.dat"],
    "forcing": ["forcing/forcing_data_resampled_uniform_Phillipsburg.csv",
                "forcing/forcing_data_resampled_uniform_Bushland.csv",
                "forcing/forcing_data_synth_0.csv", "forcing/forcing_data_synth_1.txt",
                "forcing/forcing_data_synth_2.txt"],
    "configs": ["configs/config_lasam_Phillipsburg.txt", "configs/config_lasam_Bushland.txt",
                "configs/config_lasam_Phillipsburg_with_reservoir.txt",
                "tests/configs/config_lasam_synth_0.txt", "tests/configs/config_lasam_synth_1.txt",
                "tests/configs/config_lasam_synth_2.txt", "tests/configs/unittest.txt"],
}
for sub, files in COPY.items():
    (OUT / sub).mkdir(parents=True, exist_ok=True)
    for f in files:
        shutil.copyfile(REF / f, OUT / sub / Path(f).name)
        print("copied", f)
