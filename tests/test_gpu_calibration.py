"""Calibratable parameters through SetValue (reference src/bmi_lgar.cxx:916-1078,1372-1421):
per-column soil parameter sets AND per-column reservoir / preferential-flow / ponding / field-capacity parameters (a, b,
frac_to_CR, spf_factor, ponded_depth_max, field_capacity; reference src/bmi_lgar.cxx:1032-1049,1403-1420), applied at the first
Update of a run configured with calib_params=true.  Checked against the UNMODIFIED reference (oracle/_ref, one
BmiLGAR object per parameter set) step by step: the 12 BMI scalars, the front counts, the final fronts and
volchange_calib of the global mass balance."""
import numpy as np
import pytest

from helpers import assert_fronts_close, assert_scalars_close, load_golden
from oracle.oracle_py import DATA, absolutize_config, write_config

pytestmark = pytest.mark.gpu

SOIL_SETS = [   # name -> value, per parameter set; anything not listed keeps the value Initialize() gave it
    {},
    {"smcmax_1": 0.47, "smcmax_2": 0.42, "smcmax_3": 0.40},
    {"smcmin_1": 0.03, "van_genuchten_n_2": 1.35, "hydraulic_conductivity_1": 0.9},
    {"van_genuchten_alpha_1": 0.012, "van_genuchten_alpha_3": 0.02, "hydraulic_conductivity_3": 0.05, "smcmax_2": 0.45},
    {"van_genuchten_n_1": 1.6, "van_genuchten_n_3": 1.25, "smcmin_2": 0.06, "hydraulic_conductivity_2": 0.2},
    {"smcmax_1": 0.40, "smcmin_1": 0.08, "van_genuchten_alpha_2": 0.006, "hydraulic_conductivity_1": 0.3, "smcmax_3": 0.5},
]
GLOBALS = {"a": 0.002, "b": 2.2, "frac_to_CR": 0.15, "spf_factor": 0.8, "field_capacity": 300.0, "ponded_depth_max": 0.4}
# the same parameters, a different set for every parameter set of a batch (a calibration run: one column per candidate)
GLOBAL_SETS = [
    {},
    {"a": 0.002, "b": 2.2, "frac_to_CR": 0.15, "spf_factor": 0.8, "field_capacity": 300.0, "ponded_depth_max": 0.4},
    {"a": 0.0005, "b": 3.0, "frac_to_CR": 0.35},
    {"spf_factor": 0.95, "ponded_depth_max": 1.5, "field_capacity": 250.0},
    {"a": 0.004, "b": 1.8, "frac_to_CR": 0.05, "spf_factor": 0.7, "ponded_depth_max": 0.05},
    {"b": 2.8, "field_capacity": 400.0},
]


def run_case(ref, tmp_path, cfg_keys, soil_sets, glob, T=500, reps=3, glob_sets=None):
    import torch
    import lgar_c_b200
    g = load_golden("Phillipsburg_with_reservoir")
    base = write_config(DATA / "configs" / "config_lasam_Phillipsburg_with_reservoir.txt", tmp_path / "calib.txt", calib_params="true", **cfg_keys)
    cfg = absolutize_config(base, out_dir=tmp_path)
    nset = len(soil_sets)
    ncol = nset * reps
    shifts = (np.arange(ncol) // nset) * 2500 + 40      # every repetition of a set sees another stretch of the year
    Pm = np.ascontiguousarray(np.stack([np.roll(g["precip"], -int(s))[:T] for s in shifts], axis=1))
    Em = np.ascontiguousarray(np.stack([np.roll(g["pet"], -int(s))[:T] for s in shifts], axis=1))
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol)
    names = sorted({n for s in soil_sets for n in s})
    initial = {n: b.get_value(n) for n in names}
    for n in names:
        v = initial[n].copy()
        for c in range(ncol):
            if n in soil_sets[c % nset]:
                v[c] = soil_sets[c % nset][n]
        b.set_value(n, v)
        # set / get round trip (tests/main_unit_test_bmi.cxx:575-848); the layer-3 names are accepted and ignored like in
        # the reference, whose GetVarGrid does not list them (src/bmi_lgar.cxx:1133-1141: SetValue copies 0 bytes)
        assert np.array_equal(b.get_value(n), initial[n] if n.endswith("_3") else v)
    for n, x in glob.items():
        b.set_value(n, np.full(ncol, x))
        assert np.all(b.get_value(n) == x)
    if glob_sets is not None:                              # a different reservoir / ponding set per column
        gnames = sorted({n for s in glob_sets for n in s})
        ginitial = {n: b.get_value(n) for n in gnames}
        for n in gnames:
            v = ginitial[n].copy()
            for c in range(ncol):
                if n in glob_sets[c % nset]:
                    v[c] = glob_sets[c % nset][n]
            b.set_value(n, v)
            assert np.array_equal(b.get_value(n), v)
    dP, dE = torch.from_numpy(Pm).cuda(), torch.from_numpy(Em).cuda()
    sinks = torch.zeros((12, T, ncol), dtype=torch.float64, device="cuda")
    nfs = torch.zeros((T, ncol), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    b.update_steps_device(T, dP.data_ptr(), dE.data_ptr(), ncol, {n: sinks[k].data_ptr() for k, n in enumerate(lgar_c_b200.OUTPUT_SCALARS)}, ncol, nfs.data_ptr())
    b.synchronize()
    got = sinks.cpu().numpy().transpose(1, 0, 2)          # [T][12][ncol]
    nf = nfs.cpu().numpy()
    fr = b.get_fronts()
    gm = b.get_global_mass_balance()
    assert (b.get_status() & lgar_c_b200.STATUS_FAIL_MASK == 0).all()
    for c in range(ncol):
        h = ref.create(cfg)
        for n in names:                                    # the reference holds the same initial values
            assert ref.get_scalar(h, n) == initial[n][c], n
        for n, x in soil_sets[c % nset].items():
            ref.set_scalar(h, n, x)
        for n, x in glob.items():
            ref.set_scalar(h, n, x)
        if glob_sets is not None:
            for n, x in glob_sets[c % nset].items():
                ref.set_scalar(h, n, x)
        r = ref.run_series(h, Pm[:, c], Em[:, c], cap=16)
        assert r["done"] == T, f"reference stopped at step {r['done']} for parameter set {c % nset}"
        assert np.array_equal(nf[:, c], r["nfronts"]), f"column {c}: front counts differ"
        assert_scalars_close(got[:, :, c], r["scalars"], f"column {c} (set {c % nset})")
        n_end = int(r["nfronts"][-1])
        arr = np.stack([fr[k][c] for k in ("depth", "theta", "psi", "K", "dzdt")], axis=1)
        flags = (fr["layer"][c] | (fr["to_bottom"][c] << 8)).astype(np.int32)
        assert_fronts_close(arr, flags, r["fronts"][-1], r["flags"][-1], n_end, f"column {c} final fronts")
        cum = ref.cumulative(h)
        assert abs(gm[c, 13] - cum[13]) <= 1e-12 + 1e-12 * abs(cum[13]), (c, gm[c, 13], cum[13])   # volchange_calib_cm
        ref.destroy(h)
    return gm


def test_per_column_soil_parameter_sets(cuda_device, ref, tmp_path):
    gm = run_case(ref, tmp_path, {}, SOIL_SETS, GLOBALS)
    assert np.count_nonzero(gm[:, 13]) >= 12             # the parameter update did change the stored water


def test_per_column_reservoir_and_ponding_parameter_sets(cuda_device, ref, tmp_path):
    """Every column its own (a, b, frac_to_CR, spf_factor, ponded_depth_max, field_capacity) on top of its own soil parameters:
    each against one unmodified BmiLGAR object that was given the same values through SetValue."""
    run_case(ref, tmp_path, {}, SOIL_SETS, {}, glob_sets=GLOBAL_SETS)


def test_layers_sharing_a_soil_type(cuda_device, ref, tmp_path):
    """Layers 1 and 2 of the same soil type: the reference keeps ONE row per type, so layer 1 ends up with the
    parameters given for layer 2 after its front has been recomputed with its own (src/bmi_lgar.cxx:929-990)."""
    sets = [{"smcmax_1": 0.47, "smcmax_2": 0.43, "hydraulic_conductivity_1": 0.6, "hydraulic_conductivity_2": 0.25},
            {"van_genuchten_n_1": 1.5, "smcmin_2": 0.05}]
    run_case(ref, tmp_path, {"layer_soil_type": "13,13,15"}, sets, {}, T=400, reps=2)


def test_log_mode_parameters(cuda_device, ref, tmp_path):
    """log_mode: alpha, Ksat and a are given as log10 (src/bmi_lgar.cxx:960-962,1044-1049)."""
    sets = [{"van_genuchten_alpha_1": np.log10(0.012), "van_genuchten_alpha_2": np.log10(0.008), "van_genuchten_alpha_3": np.log10(0.02),
             "hydraulic_conductivity_1": np.log10(0.5), "hydraulic_conductivity_2": np.log10(0.1), "hydraulic_conductivity_3": np.log10(0.04)}]
    run_case(ref, tmp_path, {"log_mode": "true"}, sets, {"a": np.log10(0.0015)}, T=400, reps=2)
