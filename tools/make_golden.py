#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run in the build container (needs /root/reference to (re)build oracle/_ref):
    python tools/make_golden.py
Each fixture holds the forcing used, the 12 per-step BMI scalars (metres), the per-step front
counts, and the raw wetting-front list (depth, theta, psi, K, dzdt, layer|to_bottom<<8) at a few
checkpoint steps and at the end, all in full double precision.
"""
import sys, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.oracle_py import RefHarness, read_forcing, DATA, parse_config, resolve_data_path, write_config, build_ref

OUT = ROOT / "tests" / "golden"

def run(ref, cfg, P, E, name, extra=None, every=None):
    h = ref.create(cfg)
    r = ref.run_series(h, P, E, cap=16)
    assert r["done"] == len(P), (name, r["done"])
    n = len(P)
    ck = sorted(set(list(range(0, n, every or max(1, n // 40))) + [n - 1]))
    cum = ref.cumulative(h)
    ref.destroy(h)
    d = dict(precip=P, pet=E, scalars=r["scalars"], nfronts=r["nfronts"], ck_steps=np.array(ck, dtype=np.int32),
             ck_fronts=r["fronts"][ck], ck_flags=r["flags"][ck], cumulative=cum)
    if extra: d.update(extra)
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(name, n, "steps, max fronts", int(r["nfronts"].max()))

def main():
    build_ref()
    ref = RefHarness()
    OUT.mkdir(parents=True, exist_ok=True)
    # the reference's own unit test: one Update with P=1.896, PET=0.104 mm/h (tests/main_unit_test_bmi.cxx:424-452)
    run(ref, DATA/"configs"/"unittest.txt", np.array([1.896, 0.5, 0.0, 0.0, 3.0]), np.array([0.104, 0.1, 0.2, 0.3, 0.0]), "unittest", every=1)
    for nm in ["synth_0", "synth_1", "synth_2", "Phillipsburg", "Phillipsburg_with_reservoir", "Bushland"]:
        cfg = DATA/"configs"/f"config_lasam_{nm}.txt"
        c = parse_config(cfg)
        P, E = read_forcing(resolve_data_path(c["forcing_file"][0], cfg))
        end_h = float(c["endtime"][0]) * (1.0 if c["endtime"][1] in ("[h]", "[hr]") else 1 / 3600.)
        n = int(end_h / (float(c["forcing_resolution"][0]) / 3600.))
        run(ref, cfg, P[:n], E[:n], nm)
    # random soils on the Bushland layering (BASELINE config 3) and stress / flag variants
    rng = np.random.default_rng(12345)
    PB, EB = read_forcing(DATA/"forcing"/"forcing_data_resampled_uniform_Bushland.csv")
    PP, EP = read_forcing(DATA/"forcing"/"forcing_data_resampled_uniform_Phillipsburg.csv")
    variants = [dict(), dict(timestep="300[sec]", allow_flux_caching="false", use_closed_form_G="false"),
                dict(free_drainage_enabled="true"), dict(PET_affects_precip="true", ponded_depth_max="1.5[cm]", timestep="600[sec]"),
                dict(a_slow="0.0005", b_slow="1.5", frac_slow="0.3", ponded_depth_max="0.5[cm]"),
                dict(free_drainage_enabled="true", free_drainage_to_CR="true", timestep="300[sec]")]
    tmpd = Path(tempfile.mkdtemp())
    for i in range(12):
        soils = [int(x) for x in rng.integers(1, 13, size=3)]
        var = variants[i % len(variants)]
        base = DATA/"configs"/("config_lasam_Bushland.txt" if i % 2 == 0 else "config_lasam_Phillipsburg_with_reservoir.txt")
        rep = dict(soil_params_file=str(DATA/"vG_params_stat_nom_ordered.dat"), layer_soil_type=",".join(map(str, soils)), max_valid_soil_types="12")
        rep.update(var)
        cfg = write_config(base, tmpd/f"rand_{i}.txt", **rep)
        shift = int(rng.integers(0, 8760)); mult = 0.5 + float(rng.random()) * 2.0
        Pf, Ef = (PB, EB) if i % 2 == 0 else (PP, EP)
        ns = 2500 if "timestep" not in var else 800
        P = np.roll(Pf, -shift)[:ns] * mult; E = np.roll(Ef, -shift)[:ns]
        # the fixture carries the config text with the soil file as a bare name (resolved against data/ by the tests)
        txt = cfg.read_text().replace(str(DATA) + "/", "")
        run(ref, cfg, P, E, f"rand_{i}", extra=dict(config_text=np.array(txt), base=np.array(base.name)))

if __name__ == "__main__":
    main()
