#!/usr/bin/env python
"""Random sweep: oracle vs unmodified reference over random soil triples / flag variants."""
import sys, os, tempfile
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.oracle_py import Oracle, RefHarness, read_forcing, DATA, write_config
from tools.compare_oracle_ref import compare

def main(nsets=40, seed=12345, nsteps=8333):
    rng = np.random.default_rng(seed)
    orc, ref = Oracle(), RefHarness()
    P, E = read_forcing(DATA/"forcing"/"forcing_data_resampled_uniform_Bushland.csv")
    P2, E2 = read_forcing(DATA/"forcing"/"forcing_data_resampled_uniform_Phillipsburg.csv")
    variants = [
        dict(),
        dict(timestep="300[sec]", allow_flux_caching="false", use_closed_form_G="false"),
        dict(free_drainage_enabled="true"),
        dict(free_drainage_enabled="true", free_drainage_to_CR="true", timestep="300[sec]"),
        dict(PET_affects_precip="true", ponded_depth_max="1.5[cm]", timestep="600[sec]"),
        dict(a_slow="0.0005", b_slow="1.5", frac_slow="0.3", ponded_depth_max="0.5[cm]"),
        dict(adaptive_timestep="false", timestep="1800[sec]", allow_flux_caching="false"),
    ]
    bad = 0
    tmpd = tempfile.mkdtemp()
    for i in range(nsets):
        soils = rng.integers(1, 13, size=3)
        var = variants[i % len(variants)]
        base = DATA/"configs"/("config_lasam_Bushland.txt" if i % 2 == 0 else "config_lasam_Phillipsburg_with_reservoir.txt")
        cfg = write_config(base, Path(tmpd)/f"c{i}.txt", soil_params_file=str(DATA/"vG_params_stat_nom_ordered.dat"),
                           layer_soil_type=",".join(str(s) for s in soils), max_valid_soil_types="12", **var)
        shift = int(rng.integers(0, 8760)); mult = 0.5 + rng.random()*2.0
        PP, EE = (P, E) if i % 2 == 0 else (P2, E2)
        ns = nsteps if "timestep" not in var else min(nsteps, 3000)
        f = (np.roll(PP, -shift)[:ns]*mult, np.roll(EE, -shift)[:ns])
        res, r, o = compare(cfg, forcing=f, verbose=False, orc=orc, ref=ref)
        ok = res["bit_equal_scalars"] and res["bit_equal_fronts"] and res["done_ref"] == res["done_orc"]
        print(i, list(soils), var, "OK" if ok else "MISMATCH", {k: res[k] for k in ("done_ref","done_orc","max_fronts","max_abs_scalar","theta_calls_per_step","first_front_diff","first_scalar_diff") if k in res})
        bad += (not ok)
    print("mismatches:", bad)

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 40)
