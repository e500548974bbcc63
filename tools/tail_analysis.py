#!/usr/bin/env python
"""What is the end of a scheduler window made of?  CPU analysis with the oracle's counters (no GPU): for a sample of
bench-workload columns (BASELINE config 5), after a 720 h spin-up, per column over a 240 h window: active steps A (steps that
compute fluxes instead of replaying cached ones), psi-search trips per active step, theta evaluations.

    python tools/tail_analysis.py [--columns 4000] [--window 240] [--cores 8]
Test infrastructure (executes oracle/)."""
import argparse, ctypes as C, json, multiprocessing as mp, sys, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools")); sys.path.insert(0, str(ROOT / "tests"))
import parity_sampling_full as pf  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402

_G = {}


def work(rng):
    j0, j1 = rng
    c, ids, P, E, spin = _G["c"], _G["ids"], _G["P"], _G["E"], _G["spin"]
    orc = Oracle()
    out = []
    soils, layers = c.soils(ids[j0:j1]), c.layers(ids[j0:j1])
    T = P.shape[0]
    for j in range(j0, j1):
        p = orc.params_from_config(c.cfg, layer_soil_type=[int(x) for x in soils[j - j0]], layer_thickness=[float(x) for x in layers[j - j0]])
        s = orc.new_state(p)
        orc.run_series(p, s, P[:spin, j], E[:spin, j], want_fronts=False)
        act = np.zeros(T - spin, dtype=np.int32); trips = np.zeros(T - spin, dtype=np.int64); th = np.zeros(T - spin, dtype=np.int64)
        for t in range(spin, T):
            a0, i0, h0 = s.n_active_substeps, s.n_tmb_iters, s.n_theta_calls
            orc.run_series(p, s, P[t:t + 1, j], E[t:t + 1, j], want_fronts=False)
            act[t - spin] = s.n_active_substeps - a0
            trips[t - spin] = s.n_tmb_iters - i0
            th[t - spin] = s.n_theta_calls - h0
        out.append((int((act > 0).sum()), int(trips.sum()), int(trips.max()), int(th.sum()), int(th.max())))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=4000)
    ap.add_argument("--window", type=int, default=240)
    ap.add_argument("--spinup", type=int, default=720)
    ap.add_argument("--cores", type=int, default=8)
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r1_tail_analysis.json"))
    a = ap.parse_args()
    import torch
    tmp = Path(tempfile.mkdtemp())
    c = pf.Config(5, tmp)
    ids = np.sort(np.random.default_rng(5).choice(c.n_total, a.columns, replace=False)).astype(np.int64)
    P, E = (x.numpy() for x in c.forcing(torch.from_numpy(ids), 0, a.spinup + a.window))
    _G.update(c=c, ids=ids, P=P, E=E, spin=a.spinup)
    step = max(1, a.columns // (4 * a.cores))
    jobs = [(j, min(j + step, a.columns)) for j in range(0, a.columns, step)]
    with mp.get_context("fork").Pool(a.cores) as pool:
        res = np.array([r for part in pool.map(work, jobs) for r in part], dtype=np.int64)
    A, trips, tmax, th, thmax = res.T
    K = a.window
    order = np.argsort(-A)
    ntop = max(1, int(round(0.013 * a.columns)))   # the share of columns the finishing kernel gets (16 384 of 1.25 M)
    top = order[:ntop]
    q = lambda x, pr: [float(np.percentile(x, p)) for p in pr]
    rep = dict(columns=int(a.columns), window_hours=K,
               active_fraction=float(A.sum() / (a.columns * K)),
               active_steps_per_column=dict(mean=float(A.mean()), pct=dict(zip(["p50", "p90", "p99", "p99.9", "max"], q(A, [50, 90, 99, 99.9, 100])))),
               hand_over_round_estimate=int(A[order[ntop - 1]]),
               search_trips_per_active_step=dict(all=float(trips.sum() / max(1, A.sum())), top=float(trips[top].sum() / max(1, A[top].sum()))),
               theta_evals_per_active_step=dict(all=float(th.sum() / max(1, A.sum())), top=float(th[top].sum() / max(1, A[top].sum()))),
               largest_single_step_trips=dict(all_pct=dict(zip(["p50", "p99", "max"], q(tmax, [50, 99, 100]))), top_pct=dict(zip(["p50", "p99", "max"], q(tmax[top], [50, 99, 100])))),
               top_columns=dict(count=int(ntop), active_steps_min=int(A[top].min()), active_steps_max=int(A[top].max()),
                                remaining_steps_after_hand_over_max=int(A[top].max() - A[order[ntop - 1]]),
                                trips_total_pct=dict(zip(["p50", "p90", "max"], q(trips[top], [50, 90, 100]))),
                                theta_total_pct=dict(zip(["p50", "p90", "max"], q(th[top], [50, 90, 100])))))
    Path(a.out).write_text(json.dumps(rep, indent=1))
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
