"""The plain-C oracle against (a) the reference's own known answers and (b) golden vectors generated
from the unmodified reference (tools/make_golden.py).  CPU only.  The oracle runs the reference's
arithmetic on the same libm, so the bar here is bit equality, not 1e-9."""
import numpy as np
import pytest

from helpers import STANDARD, RANDOM, load_golden, config_for, assert_physics_sanity


@pytest.mark.parametrize("name", STANDARD + RANDOM)
def test_oracle_matches_reference_golden(oracle, name, tmp_path):
    g = load_golden(name)
    cfg = config_for(name, tmp_path)
    p = oracle.params_from_config(cfg)
    s = oracle.new_state(p)
    r = oracle.run_series(p, s, g["precip"], g["pet"], cap=16)
    assert r["done"] == len(g["precip"]) and s.status & 0xFFFF == 0
    assert np.array_equal(r["nfronts"], g["nfronts"])
    assert np.array_equal(r["scalars"], g["scalars"]), "per-step BMI scalars are not bit-identical to the reference"
    ck = g["ck_steps"]
    assert np.array_equal(r["fronts"][ck], g["ck_fronts"])
    assert np.array_equal(r["flags"][ck], g["ck_flags"])
    assert oracle.lib.lgo_global_balance(s) == g["cumulative"][15]


@pytest.mark.parametrize("name", ["synth_1", "synth_2"])
def test_synthetic_cases_against_hydrus_and_lgar_py(oracle, name, tmp_path):
    """BASELINE config 2 asks for the synthetic cases to be checked against the HYDRUS / Python outputs the reference ships
    (tests/HYDRUS_outputs, tests/Python_outputs): soil water storage and cumulative runoff of the oracle AND of the golden
    vector of the unmodified reference within the millimetre-level distances SURVEY.md section 4 measured."""
    g = load_golden(name)
    p = oracle.params_from_config(config_for(name, tmp_path))
    r = oracle.run_series(p, oracle.new_state(p), g["precip"], g["pet"], cap=16)
    d = assert_physics_sanity(name, r["scalars"][:, 5], r["scalars"][:, 3])
    assert d == assert_physics_sanity(name, g["scalars"][:, 5], g["scalars"][:, 3])


def test_reference_unit_test_known_answers(oracle, tmp_path):
    """tests/main_unit_test_bmi.cxx:456-457,504-522 of the reference: one Update() with P=1.896 mm/h,
    PET=0.104 mm/h must give 4 fronts at these depths / moistures (1e-5) and these fluxes."""
    cfg = config_for("unittest", tmp_path)
    p = oracle.params_from_config(cfg)
    s = oracle.new_state(p)
    st, out = oracle.update(p, s, 1.896, 0.104)
    assert st == 0 and s.n == 4
    depth_b = [4.55355239489608365, 44.00, 175.0, 200.0]
    theta_b = [0.21371581122514613, 0.17270389607163267, 0.25211383152603861, 0.17959348005962811]
    for i in range(4):
        assert abs(s.f[i].depth - depth_b[i]) < 1e-5
        assert abs(s.f[i].theta - theta_b[i]) < 1e-5
    assert abs(out[7] * 1000 - 1.896) < 1e-5            # infiltration, mm
    assert abs(out[2] * 1000 - 0.02980092620558239) < 1e-5  # AET, mm
    assert abs(out[1] * 1000 - 0.104) < 1e-5            # PET, mm


@pytest.mark.parametrize("name,totals", [
    # BASELINE.md section 2, year totals printed by the reference's own lasam_standalone (cm)
    ("Phillipsburg", dict(volin=86.1385549628, volrunoff=12.9468450372, volAET=85.4399004917, bal=1.345e-7)),
    ("Phillipsburg_with_reservoir", dict(volin=78.7250483600, volrunoff=18.7804716400, volAET=78.2471105283, bal=1.41e-7)),
    ("Bushland", dict(volin=18.3634973403, volrunoff=7.0873026597, volAET=29.7913954767, bal=1.58e-8)),
])
def test_year_totals_match_standalone(oracle, name, totals, tmp_path):
    g = load_golden(name)
    p = oracle.params_from_config(config_for(name, tmp_path))
    s = oracle.new_state(p)
    oracle.run_series(p, s, g["precip"], g["pet"], want_fronts=False)
    assert abs(s.volin_cm - totals["volin"]) < 1e-9
    assert abs(s.volrunoff_cm - totals["volrunoff"]) < 1e-9
    assert abs(s.volAET_cm - totals["volAET"]) < 1e-9
    assert abs(oracle.lib.lgo_global_balance(s) - totals["bal"]) < 0.01 * abs(totals["bal"]) + 1e-10


def test_soil_parser_quirk_rows_with_blank_in_name(oracle):
    """rows "Sandy-clay loam"/"Silty-clay loam" inherit the previous row's numbers."""
    from oracle.oracle_py import DATA
    table, n = oracle.read_soil_file(DATA / "vG_params_stat_nom_ordered.dat", 12)
    assert n == 12
    for bad in (7, 8):
        assert (table[bad].theta_r, table[bad].theta_e, table[bad].alpha, table[bad].n, table[bad].Ksat) == (0.06, 0.4, 0.01, 1.47, 0.504)
    t2, _ = oracle.read_soil_file(DATA / "vG_default_params.dat", 25)
    assert t2[7].Ksat == t2[6].Ksat and t2[11].Ksat == t2[10].Ksat


def test_invalid_soil_type_passes_precip_through(oracle, tmp_path):
    from oracle.oracle_py import write_config, DATA
    cfg = write_config(DATA / "configs" / "config_lasam_Phillipsburg.txt", tmp_path / "inv.txt", layer_soil_type="13,26,15",
                       soil_params_file=str(DATA / "vG_default_params.dat"))
    p = oracle.params_from_config(cfg)
    assert p.invalid_soil_type == 1
    s = oracle.new_state(p)
    st, out = oracle.update(p, s, 3.0, 0.2)
    assert st == 0 and out[0] == out[3] == out[6] == 3.0 * 0.001 and out[7] == 0.0
