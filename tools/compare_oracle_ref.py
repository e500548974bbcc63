#!/usr/bin/env python
"""Step-by-step comparison of the plain-C oracle against the unmodified reference (oracle/_ref)."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.oracle_py import Oracle, RefHarness, read_forcing, DATA, parse_config, resolve_data_path

def compare(cfg, nsteps=None, verbose=True, forcing=None, orc=None, ref=None, cfgfile_for_ref=None):
    orc = orc or Oracle(); ref = ref or RefHarness()
    c = parse_config(cfg)
    if forcing is None:
        P, E = read_forcing(resolve_data_path(c["forcing_file"][0], cfg))
    else:
        P, E = forcing
    if nsteps: P, E = P[:nsteps], E[:nsteps]
    p = orc.params_from_config(cfg)
    s = orc.new_state(p)
    h = ref.create(cfgfile_for_ref or cfg)
    t0=time.time(); r = ref.run_series(h, P, E); t1=time.time()
    o = orc.run_series(p, s, P, E); t2=time.time()
    n = min(r["done"], o["done"])
    res = dict(done_ref=r["done"], done_orc=o["done"], n=len(P))
    nf_ok = np.array_equal(r["nfronts"][:n], o["nfronts"][:n])
    res["nfronts_equal"] = nf_ok
    res["first_nf_diff"] = int(np.argmax(r["nfronts"][:n] != o["nfronts"][:n])) if not nf_ok else -1
    d = np.abs(r["scalars"][:n] - o["scalars"][:n])
    res["max_abs_scalar"] = float(d.max()) if n else 0.0
    res["bit_equal_scalars"] = bool(np.array_equal(r["scalars"][:n], o["scalars"][:n]))
    res["bit_equal_fronts"] = bool(np.array_equal(r["fronts"][:n], o["fronts"][:n]) and np.array_equal(r["flags"][:n], o["flags"][:n]))
    if not res["bit_equal_scalars"]:
        bad = np.argwhere(r["scalars"][:n] != o["scalars"][:n])
        res["first_scalar_diff"] = (int(bad[0][0]), int(bad[0][1]))
    if not res["bit_equal_fronts"]:
        bad = np.argwhere((r["fronts"][:n] != o["fronts"][:n]).any(axis=(1,2)) | (r["flags"][:n] != o["flags"][:n]).any(axis=1))
        res["first_front_diff"] = int(bad[0][0])
    res["t_ref"]=round(t1-t0,3); res["t_orc"]=round(t2-t1,3)
    res["max_fronts"]=int(r["nfronts"][:n].max()) if n else 0
    cum = ref.cumulative(h)
    res["ref_global_balance"] = cum[15]
    res["orc_global_balance"] = orc.lib.lgo_global_balance(s)
    res["theta_calls_per_step"] = s.n_theta_calls/max(n,1)
    res["active_substeps"] = s.n_active_substeps
    ref.destroy(h)
    if verbose: print(Path(cfg).name, res)
    return res, r, o

if __name__ == "__main__":
    names = sys.argv[1:] or ["unittest.txt","config_lasam_synth_0.txt","config_lasam_synth_1.txt","config_lasam_synth_2.txt",
             "config_lasam_Phillipsburg.txt","config_lasam_Phillipsburg_with_reservoir.txt","config_lasam_Bushland.txt"]
    for nm in names:
        cfg = DATA/"configs"/nm
        c = parse_config(cfg)
        if nm == "unittest.txt":
            compare(cfg, forcing=(np.array([1.896, 0.5, 0.0, 0.0, 3.0]), np.array([0.104, 0.1, 0.2, 0.3, 0.0])))
            continue
        end_h = float(c["endtime"][0]) * (1.0 if c["endtime"][1] in ("[h]","[hr]") else 1/3600.)
        res_h = float(c["forcing_resolution"][0])/3600.
        compare(cfg, nsteps=int(end_h/res_h))
