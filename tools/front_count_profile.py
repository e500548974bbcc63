#!/usr/bin/env python
"""How many fronts does an ACTIVE column-step carry?  Sizing data for a compact on-chip lane record (DESIGN.md section 10,
item 1): the lane record has room for 32 fronts (3.3 KB) because 0.2 % of the columns of a one-year run need more than 16,
but what a stage kernel could keep in shared memory is set by the common case.  CPU analysis with the oracle (no GPU): a
sample of bench-workload columns (BASELINE config 5), 720 h spin-up, then per active step of a window the front count
before and after the step.

    python tools/front_count_profile.py [--columns 2000] [--window 480] [--cores 8]
Test infrastructure (executes oracle/)."""
import argparse, json, multiprocessing as mp, sys, tempfile
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
    soils, layers = c.soils(ids[j0:j1]), c.layers(ids[j0:j1])
    T = P.shape[0]
    hist_active = np.zeros(40, dtype=np.int64)   # max(front count before, after) of an active step
    hist_all = np.zeros(40, dtype=np.int64)      # front count after every step
    colmax = []
    for j in range(j0, j1):
        p = orc.params_from_config(c.cfg, layer_soil_type=[int(x) for x in soils[j - j0]], layer_thickness=[float(x) for x in layers[j - j0]])
        s = orc.new_state(p)
        r = orc.run_series(p, s, P[:spin, j], E[:spin, j], want_fronts=False)
        prev = int(r["nfronts"][-1])
        mx = 0
        for t in range(spin, T):
            a0 = s.n_active_substeps
            r = orc.run_series(p, s, P[t:t + 1, j], E[t:t + 1, j], want_fronts=False)
            if r["done"] < 1:
                break
            n = int(r["nfronts"][0])
            hist_all[min(n, 39)] += 1
            if s.n_active_substeps > a0:
                hist_active[min(max(n, prev), 39)] += 1
            mx = max(mx, n)
            prev = n
        colmax.append(mx)
    return hist_active, hist_all, colmax


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=2000)
    ap.add_argument("--window", type=int, default=480)
    ap.add_argument("--spinup", type=int, default=720)
    ap.add_argument("--cores", type=int, default=8)
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r1_front_count_profile.json"))
    a = ap.parse_args()
    import torch
    c = pf.Config(5, Path(tempfile.mkdtemp()))
    ids = np.sort(np.random.default_rng(6).choice(c.n_total, a.columns, replace=False)).astype(np.int64)
    P, E = (x.numpy() for x in c.forcing(torch.from_numpy(ids), 0, a.spinup + a.window))
    _G.update(c=c, ids=ids, P=P, E=E, spin=a.spinup)
    step = max(1, a.columns // (4 * a.cores))
    jobs = [(j, min(j + step, a.columns)) for j in range(0, a.columns, step)]
    with mp.get_context("fork").Pool(a.cores) as pool:
        parts = pool.map(work, jobs)
    ha = sum(p[0] for p in parts); hl = sum(p[1] for p in parts); cm = np.array([x for p in parts for x in p[2]])
    cdf = lambda h: {f"<= {k}": float(h[:k + 1].sum() / max(1, h.sum())) for k in (3, 4, 5, 6, 8, 10, 12, 16)}
    # LgLane (lgar_lanes.cuh) with front capacity w: 352 B header + scalars + epilogue state, the five f64 front arrays with
    # layer/to_bottom bytes and the count (padded to 32), 544 B step locals, state_previous (3 arrays + 32), 192 B search
    # record; w = 32 gives the 3 264 B of the built record
    rec = lambda w: 352 + (40 * w + 2 * w + 4 + 31) // 32 * 32 + 544 + 24 * w + 32 + 192
    rep = dict(columns=int(a.columns), window_hours=int(a.window), spinup_hours=int(a.spinup),
               active_steps=int(ha.sum()), active_fraction=float(ha.sum() / max(1, hl.sum())),
               fronts_of_an_active_step=dict(mean=float((ha * np.arange(40)).sum() / max(1, ha.sum())), cdf=cdf(ha),
                                             histogram={str(k): int(v) for k, v in enumerate(ha) if v}),
               fronts_after_any_step=dict(mean=float((hl * np.arange(40)).sum() / max(1, hl.sum())), cdf=cdf(hl)),
               column_maximum_over_window=dict(cdf={f"<= {k}": float((cm <= k).mean()) for k in (4, 6, 8, 10, 12, 16)}, max=int(cm.max())),
               record_bytes_at_capacity={str(w): int(rec(w)) for w in (6, 8, 12, 16, 32)},
               records_per_sm_in_227KB_shared={str(w): int(227 * 1024 // rec(w)) for w in (6, 8, 12, 16, 32)})
    Path(a.out).write_text(json.dumps(rep, indent=1))
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
