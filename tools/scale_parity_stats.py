#!/usr/bin/env python
"""How far does the engine deviate from the oracle at scale, and where?  (bench workload, sampled columns)"""
import sys, tempfile
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench, lgar_c_b200
from oracle.oracle_py import Oracle, SCALAR_NAMES
ncol, T, nsample = 100_000, 600, int(sys.argv[1]) if len(sys.argv) > 1 else 3000
tmp = Path(tempfile.mkdtemp()); cfg = bench.write_workload_config(tmp)
soils, layers = bench.column_soils(0, ncol), bench.column_layers(0, ncol)
P, E = bench.host_forcing(ncol, 0, T, 4242)
rng = np.random.default_rng(7); storm = rng.random((T, ncol)) < 0.004
for k in range(1, 5):
    P[k:] = np.where(storm[:-k], np.maximum(P[k:], 6.0 * rng.random((T - k, ncol))), P[k:])
b = lgar_c_b200.BmiLGARBatch(); b.initialize(cfg, ncol, 0, soils, layers)
dP, dE = torch.from_numpy(P).cuda(), torch.from_numpy(E).cuda()
sample = np.sort(np.random.default_rng(99).choice(ncol, nsample, replace=False))
sinks = torch.zeros((12, T, ncol), dtype=torch.float64, device="cuda"); nfs = torch.zeros((T, ncol), dtype=torch.int32, device="cuda")
b.update_steps_device(T, dP.data_ptr(), dE.data_ptr(), ncol, {n: sinks[k].data_ptr() for k, n in enumerate(lgar_c_b200.OUTPUT_SCALARS)}, ncol, nfs.data_ptr())
b.synchronize()
status = b.get_status(); failed = np.nonzero(status & 0xFFFF)[0]
cols = np.unique(np.concatenate([sample, failed[:500]]))
idx = torch.from_numpy(cols).cuda()
scal = sinks[:, :, idx].cpu().numpy().transpose(1, 0, 2); nf = nfs[:, idx].cpu().numpy()
orc = Oracle()
worst_rel = np.zeros(12); worst_abs = np.zeros(12); nf_mismatch = 0; viol = np.zeros(12, dtype=int); fail_mismatch = 0
big = []
for j, c in enumerate(cols):
    p = orc.params_from_config(cfg, layer_soil_type=[int(x) for x in soils[c]], layer_thickness=[float(x) for x in layers[c]])
    s = orc.new_state(p); r = orc.run_series(p, s, P[:, c], E[:, c], want_fronts=False); done = r["done"]
    gfail = bool(status[c] & 0xFFFF)
    if gfail != (done < T) or (gfail and (s.status & 0xFFFF) != (status[c] & 0xFFFF)): fail_mismatch += 1
    if not np.array_equal(nf[:done, j], r["nfronts"][:done]): nf_mismatch += 1; big.append((int(c), "nfronts", int(np.argmax(nf[:done, j] != r["nfronts"][:done]))))
    g, w = scal[:done, :, j], r["scalars"][:done]
    err = np.abs(g - w); rel = err / np.maximum(np.maximum(np.abs(g), np.abs(w)), 1e-300)
    relm = np.where(err > 1e-14, rel, 0.0)
    worst_rel = np.maximum(worst_rel, np.nan_to_num(relm).max(axis=0)); worst_abs = np.maximum(worst_abs, np.nan_to_num(err).max(axis=0))
    v = (err > 1e-9 * np.maximum(np.abs(g), np.abs(w)) + 1e-14)
    viol += v.any(axis=0)
print("columns checked", len(cols), "failed on GPU", len(failed), "of", ncol, "| failure mismatches vs oracle:", fail_mismatch, "| front-count mismatches:", nf_mismatch)
for k, n in enumerate(SCALAR_NAMES):
    print(f"  {n:42s} worst rel {worst_rel[k]:.2e}  worst abs {worst_abs[k]:.2e} m   columns violating 1e-9 rel: {viol[k]}")
print(big[:10])
