#!/usr/bin/env python
"""Which of the reference's rare branches do the BASELINE workloads reach, and in which column first?  The oracle's path counters
(oracle/lgar_oracle.h LGO_PATH_*) over 2 000 columns of the seed-99 sample of configs 3, 4 and 5 for their full duration.  This is
where tests/test_gpu_scale_parity.py::RARE_PATH_COLUMNS comes from.  Result of the committed run: profiles/r2_path_scan.txt --
merge, layer crossing, domain-bottom crossing, dry-over-wet, theta < theta_r deletion (configs 3 and 5), saturated-front depth loop,
crossing + bottom depth search are all reached; lgar_clean_redundant_fronts never deletes a front and no psi search ends on
anything but convergence in 5e7 column-steps (the golden synthetic cases have three searches that end on the step-size exit).

    python tools/path_scan.py        (test infrastructure: executes oracle/)"""
import multiprocessing as mp
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tools"), str(ROOT / "tests")]
import parity_sampling_full as pf  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402

NAMES = ["merge", "cross_layer", "cross_bottom", "dry_over_wet", "theta_r_delete", "saturated_depth", "cross_layer_bottom", "surficial",
         "clean_redundant", "search", "search_converged", "search_step_exit", "search_nochange", "search_maxiter", "search_psi_upper",
         "search_psi_zero", "corr_search", "cached_step", "active_step"]


def work(a):
    import torch
    cfgn, j0, j1, T = a
    c = pf.Config(cfgn, Path(tempfile.mkdtemp()))
    ids = np.sort(np.random.default_rng(99).choice(c.n_total, 2000, replace=False)).astype(np.int64)[j0:j1]
    P, E = (x.numpy() for x in c.forcing(torch.from_numpy(ids), 0, T))
    orc = Oracle()
    soils, layers = c.soils(ids), c.layers(ids)
    tot, first = np.zeros(24, dtype=np.int64), {}
    for j in range(len(ids)):
        kw = {}
        if soils is not None:
            kw["layer_soil_type"] = [int(x) for x in soils[j]]
        if layers is not None:
            kw["layer_thickness"] = [float(x) for x in layers[j]]
        p = orc.params_from_config(c.cfg, **kw)
        s = orc.new_state(p)
        orc.run_series(p, s, np.ascontiguousarray(P[:, j]), np.ascontiguousarray(E[:, j]), want_fronts=False)
        v = np.array(s.n_path[:])
        tot += v
        for i in np.flatnonzero(v):
            first.setdefault(int(i), int(ids[j]))
    return tot, first


if __name__ == "__main__":
    for cfgn, T in ((5, 8760), (4, 7500), (3, 8333)):
        with mp.get_context("fork").Pool(8) as pool:
            res = pool.map(work, [(cfgn, j, j + 250, T) for j in range(0, 2000, 250)])
        tot, first = sum(r[0] for r in res), {}
        for r in res:
            for k, v in r[1].items():
                first.setdefault(k, v)
        print(f"config {cfgn}: 2000 columns x {T} steps")
        for i, nm in enumerate(NAMES):
            print("  %-20s %10d  first column %s" % (nm, tot[i], first.get(i)))
