#!/usr/bin/env python
"""tests/golden/physics_synth_{1,2}.npz: the independent solutions the reference validates its synthetic cases against
(tests/plot_synthetic_examples.ipynb: HYDRUS-1D T_Level.txt and the LGAR-Py pickle), reduced to the series the notebook
overlays: time [min], soil water storage [mm], cumulative runoff / infiltration [mm].

Run in the build container (reads /root/reference/tests/{HYDRUS,Python}_outputs; those files do not travel to the GPU box):
    python tools/make_physics_golden.py
"""
from pathlib import Path

import numpy as np
import pandas as pd

REF = Path("/root/reference/tests")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def main():
    for k in (1, 2):
        # HYDRUS-1D: fixed-width columns of 13 characters (the notebook reads it with pd.read_fwf(widths=[13]*22)); times in minutes
        rows = []
        for line in (REF / "HYDRUS_outputs" / f"synth_{k}" / "T_Level.txt").read_text().splitlines()[1:]:
            f = line.split()
            if len(f) >= 19:
                rows.append([float(x) for x in f[:19]])
        h = np.array(rows)
        py = pd.read_pickle(REF / "Python_outputs" / f"synth_{k}" / f"output_synth_{k}.pkl")
        dt_h = (py.index[1] - py.index[0]).total_seconds() / 3600.0
        np.savez_compressed(
            OUT / f"physics_synth_{k}.npz",
            hydrus_time_min=h[:, 0], hydrus_storage_mm=h[:, 16], hydrus_cum_runoff_mm=h[:, 15], hydrus_cum_infil_mm=h[:, 17],
            py_storage_mm=py["water_in_soil[mm]"].to_numpy(float), py_runoff_mm=py["runoff[mm/h]"].to_numpy(float) * dt_h,
            py_infil_mm=py["actual_infil[mm/h]"].to_numpy(float) * dt_h, py_dt_h=np.float64(dt_h))
        print(k, h.shape, "hydrus storage", h[0, 16], "->", h[-1, 16], "cum runoff", h[-1, 15], "| py steps", len(py), "dt_h", dt_h)


if __name__ == "__main__":
    main()
