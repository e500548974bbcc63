"""Parity of the CUDA engine (through the C ABI) with the oracle / the reference's golden vectors.
Bar (BASELINE.json north_star): every per-step result within 1e-9 relative, front counts exact.  The tests ask for more: the
engine's pow is the host library's (lgar_pow.cuh), so every compared value must equal the reference's bit for bit (helpers.EXACT).
All tests here need a B200 (pytest -m gpu); they never read /root/reference."""
import numpy as np
import pytest

from helpers import STANDARD, RANDOM, load_golden, config_for, assert_scalars_close, assert_fronts_close, RTOL

pytestmark = pytest.mark.gpu


def engine_run(cfg, P, E, ncol=3, soil_types=None, thickness=None, force_all_active=False, shifts=None, mult=None):
    """Run the engine over a [T] forcing series (optionally per-column circularly shifted / scaled)
    with the forcing resident on the device; returns per-step scalars [T][12][ncol], nfronts [T][ncol]
    and the engine (for final-state queries)."""
    import torch
    import lgar_c_b200
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol, 0, soil_types, thickness)
    if force_all_active:
        b.set_force_all_active(True)
    T = len(P)
    Pm = np.empty((T, ncol)); Em = np.empty((T, ncol))
    for c in range(ncol):
        sh = 0 if shifts is None else int(shifts[c])
        m = 1.0 if mult is None else float(mult[c])
        Pm[:, c] = np.roll(P, -sh)[:T] * m
        Em[:, c] = np.roll(E, -sh)[:T]
    dev = torch.device("cuda:0")
    dP = torch.from_numpy(Pm).to(dev); dE = torch.from_numpy(Em).to(dev)
    sinks = torch.full((12, T, ncol), float("nan"), dtype=torch.float64, device=dev)
    nfs = torch.zeros((T, ncol), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    b.update_steps_device(T, dP.data_ptr(), dE.data_ptr(), ncol,
                          {n: sinks[k].data_ptr() for k, n in enumerate(lgar_c_b200.OUTPUT_SCALARS)}, ncol, nfs.data_ptr())
    b.synchronize()
    return b, sinks.cpu().numpy().transpose(1, 0, 2), nfs.cpu().numpy(), (Pm, Em)


def fronts_record(fr, c):
    n = int(fr["nfronts"][c])
    arr = np.stack([fr[k][c] for k in ("depth", "theta", "psi", "K", "dzdt")], axis=1)
    flags = (fr["layer"][c] | (fr["to_bottom"][c] << 8)).astype(np.int32)
    return arr, flags, n


@pytest.mark.parametrize("name", STANDARD + RANDOM)
def test_engine_matches_reference_golden(cuda_device, name, tmp_path):
    g = load_golden(name)
    cfg = config_for(name, tmp_path)
    T = len(g["precip"])
    b, scal, nf, _ = engine_run(cfg, g["precip"], g["pet"], ncol=3)
    st = b.get_status()
    assert (st & 0xFFFF == 0).all(), st
    for c in range(3):
        assert np.array_equal(nf[:, c], g["nfronts"]), f"{name}: front count differs at step {int(np.argmax(nf[:, c] != g['nfronts']))}"
        assert_scalars_close(scal[:, :, c], g["scalars"], f"{name} col {c}")
    fr = b.get_fronts()
    for c in range(3):
        arr, flags, n = fronts_record(fr, c)
        assert n == g["nfronts"][-1]
        assert_fronts_close(arr, flags, g["ck_fronts"][-1], g["ck_flags"][-1], n, f"{name} final fronts")
    mb = b.get_global_mass_balance()
    ref_cum = g["cumulative"]
    # volprecip, volin, volrunoff, volAET, volrech, volQ: the cumulative sums of the global balance, same additions in the same order
    for mine, theirs in ((1, 1), (2, 2), (5, 5), (7, 7), (8, 8), (9, 9), (10, 10), (12, 12)):
        assert mb[0, mine] == ref_cum[theirs], (mine, mb[0, mine], ref_cum[theirs])
    assert abs(mb[0, 15] - ref_cum[15]) < 1e-8
    b.finalize()


def test_reference_unit_test_known_answers_on_gpu(cuda_device, tmp_path):
    """reference tests/main_unit_test_bmi.cxx:424-568 through the BMI verbs of the C ABI"""
    import lgar_c_b200
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(config_for("unittest", tmp_path), 2)
    b.set_value("precipitation_rate", [1.896, 1.896])
    b.set_value("potential_evapotranspiration_rate", [0.104, 0.104])
    assert list(b.get_value("precipitation_rate")) == [1.896, 1.896]
    b.update()
    n = b.get_value("soil_num_wetting_fronts")
    assert list(n) == [4, 4]
    depth = b.get_value("soil_depth_wetting_fronts")
    theta = b.get_value("soil_moisture_wetting_fronts")
    depth_b = [4.55355239489608365, 44.00, 175.0, 200.0]
    theta_b = [0.21371581122514613, 0.17270389607163267, 0.25211383152603861, 0.17959348005962811]
    for c in range(2):
        for i in range(4):
            assert abs(depth[c, i] * 100 - depth_b[i]) < 1e-5 and abs(theta[c, i] - theta_b[i]) < 1e-5
        assert np.isnan(depth[c, 4:]).all()
    assert abs(b.get_value("infiltration")[0] * 1000 - 1.896) < 1e-5
    assert abs(b.get_value("actual_evapotranspiration")[0] * 1000 - 0.02980092620558239) < 1e-5
    assert abs(b.get_value("potential_evapotranspiration")[0] * 1000 - 0.104) < 1e-5
    assert b.get_current_time() == 3600.0 and b.get_time_step() == 3600.0 and b.get_end_time() == 3600.0


def test_device_pow_equals_host_pow(cuda_device, tmp_path):
    """lgar_c_b200/csrc/lgar_pow.cuh ON THE DEVICE against the pow of the host's C library -- the function the reference calls
    (reference src/soil_funcs.cxx:147-187) -- bit for bit: 16 M pairs of the shapes the path produces, the lattice of special
    values (zeros, infinities, NaN, negative bases with integer and non-integer exponents, subnormals) and results in the
    subnormal and overflow ranges; both the out-of-line function and the inlined common case of the psi search."""
    import ctypes as C
    import lgar_c_b200
    from test_pow_exact import path_samples, special_samples, edge_samples, same_bits
    libm = C.CDLL("libm.so.6")
    libm.pow.restype = C.c_double
    libm.pow.argtypes = [C.c_double, C.c_double]
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(config_for("Phillipsburg", tmp_path), 1)
    xs, ys = zip(path_samples(2_000_000, 11), special_samples(), edge_samples(500_000, 12))
    x, y = np.concatenate(xs), np.concatenate(ys)
    import math
    # numpy's power is not libm's pow on every build: take the host value from libm itself on a subsample, numpy on the rest
    want = np.power(x, y)
    idx = np.random.default_rng(3).choice(len(x), 200_000, replace=False)
    host = np.array([libm.pow(float(x[i]), float(y[i])) for i in idx])
    for which in (0, 1):
        got = b.debug_pow(x, y, which)
        bad = ~same_bits(got[idx], host)
        assert not bad.any(), [(float(x[idx][i]).hex(), float(y[idx][i]).hex(), float(got[idx][i]).hex(), float(host[i]).hex()) for i in np.flatnonzero(bad)[:5]]
        if same_bits(want[idx], host).all():   # numpy agrees with libm on the subsample: use it for all 17 M pairs
            bad = ~same_bits(got, want)
            assert not bad.any(), [(float(x[i]).hex(), float(y[i]).hex(), float(got[i]).hex(), float(want[i]).hex()) for i in np.flatnonzero(bad)[:5]]
    b.finalize()


def test_two_kernel_split_equals_single_path(cuda_device, tmp_path):
    """The streaming kernel (cached-flux steps) + active kernel must be indistinguishable from
    routing every column through the active kernel."""
    g = load_golden("Phillipsburg_with_reservoir")
    cfg = config_for("Phillipsburg_with_reservoir", tmp_path)
    T = 2500
    b1, s1, n1, _ = engine_run(cfg, g["precip"][:T], g["pet"][:T], ncol=5)
    b2, s2, n2, _ = engine_run(cfg, g["precip"][:T], g["pet"][:T], ncol=5, force_all_active=True)
    assert np.array_equal(n1, n2) and np.array_equal(s1, s2)
    f1, f2 = b1.get_fronts(), b2.get_fronts()
    for k in ("depth", "theta", "psi", "K", "dzdt", "layer", "to_bottom", "nfronts"):
        assert np.array_equal(f1[k], f2[k], equal_nan=True)


def test_window_kernels_equal_hourly_kernels(cuda_device, tmp_path):
    """lgarb_update_steps_device in window mode (one launch, per-column time pointers) vs one kernel pair
    per hour: identical per-step outputs, front counts and final state."""
    import torch
    import lgar_c_b200
    g = load_golden("Phillipsburg_with_reservoir")
    cfg = config_for("Phillipsburg_with_reservoir", tmp_path)
    ncol, T = 700, 1500                      # 700 columns: more than two warp tiles, last one ragged
    shifts = (np.arange(ncol) * 7919) % 8760
    Pm = np.ascontiguousarray(np.stack([np.roll(g["precip"], -int(s))[:T] for s in shifts], axis=1))
    Em = np.ascontiguousarray(np.stack([np.roll(g["pet"], -int(s))[:T] for s in shifts], axis=1))
    res = []
    for window in (3, 1, 0):
        b = lgar_c_b200.BmiLGARBatch()
        b.initialize(cfg, ncol)
        b.set_window_mode(window)
        dP, dE = torch.from_numpy(Pm).cuda(), torch.from_numpy(Em).cuda()
        sinks = torch.zeros((12, T, ncol), dtype=torch.float64, device="cuda")
        nfs = torch.zeros((T, ncol), dtype=torch.int32, device="cuda")
        # two consecutive windows, to cross a window boundary
        for t0, n in ((0, 900), (900, T - 900)):
            torch.cuda.synchronize()  # the engine runs on its own stream: inputs/sinks made by torch must be complete
            b.update_steps_device(n, dP[t0].data_ptr(), dE[t0].data_ptr(), ncol,
                                  {nm: sinks[k, t0].data_ptr() for k, nm in enumerate(lgar_c_b200.OUTPUT_SCALARS)}, ncol, nfs[t0].data_ptr())
        b.synchronize()
        res.append((sinks.cpu().numpy(), nfs.cpu().numpy(), b.get_fronts(), b.get_global_mass_balance(), b.get_value("soil_storage"),
                    b.get_current_time()))
    (s2, n2, f2, m2, v2, t2) = res[-1]
    for (s1, n1, f1, m1, v1, t1) in res[:-1]:
        assert np.array_equal(n1, n2) and np.array_equal(s1, s2) and np.array_equal(m1, m2) and np.array_equal(v1, v2) and t1 == t2
        for k in ("depth", "theta", "psi", "K", "dzdt", "layer", "to_bottom", "nfronts"):
            assert np.array_equal(f1[k], f2[k], equal_nan=True)


def test_host_buffer_multi_step_call(cuda_device, tmp_path):
    """lgarb_update_steps_host (forcing in / output series out through host buffers) equals the device path."""
    import torch
    import lgar_c_b200
    g = load_golden("Bushland")
    cfg = config_for("Bushland", tmp_path)
    ncol, T = 300, 700
    shifts = (np.arange(ncol) * 7919) % 8760
    Pm = np.ascontiguousarray(np.stack([np.roll(g["precip"], -int(s))[:T] for s in shifts], axis=1))
    Em = np.ascontiguousarray(np.stack([np.roll(g["pet"], -int(s))[:T] for s in shifts], axis=1))
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol)
    out = b.update_steps_host(Pm, Em, ["total_discharge", "soil_storage", "actual_evapotranspiration"])
    b2 = lgar_c_b200.BmiLGARBatch()
    b2.initialize(cfg, ncol)
    dP, dE = torch.from_numpy(Pm).cuda(), torch.from_numpy(Em).cuda()
    sinks = torch.zeros((12, T, ncol), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()  # the engine runs on its own stream: inputs/sinks made by torch must be complete
    b2.update_steps_device(T, dP.data_ptr(), dE.data_ptr(), ncol, {n: sinks[k].data_ptr() for k, n in enumerate(lgar_c_b200.OUTPUT_SCALARS)}, ncol)
    b2.synchronize()
    ref = sinks.cpu().numpy()
    for name, k in (("total_discharge", 6), ("soil_storage", 5), ("actual_evapotranspiration", 2)):
        assert np.array_equal(out[name], ref[k]), name
    with pytest.raises(lgar_c_b200.LgarbError):
        b.update_steps_host(Pm, Em, ["soil_moisture_wetting_fronts"])


def test_streamed_host_call_equals_single_call(cuda_device, tmp_path):
    """lgarb_update_steps_host_stream (chunks, copies overlapped with compute): f64 in / f64 out is bit-identical to
    lgarb_update_steps_host whatever the chunk length (ragged last chunk included); f32 forcing equals an f64 run fed
    the same rounded values, and f32 outputs are the f64 outputs rounded once."""
    import lgar_c_b200
    g = load_golden("Phillipsburg_with_reservoir")
    cfg = config_for("Phillipsburg_with_reservoir", tmp_path)
    ncol, T = 257, 530
    shifts = (np.arange(ncol) * 7919) % 8760
    Pm = np.ascontiguousarray(np.stack([np.roll(g["precip"], -int(s))[:T] for s in shifts], axis=1))
    Em = np.ascontiguousarray(np.stack([np.roll(g["pet"], -int(s))[:T] for s in shifts], axis=1))
    names = ["total_discharge", "soil_storage", "infiltration"]

    def run(P, E, chunk=None, out_dtype=np.float64, env=None):
        import os
        os.environ.update(env or {})
        try:
            b = lgar_c_b200.BmiLGARBatch()
        finally:
            for k in (env or {}):
                del os.environ[k]
        b.initialize(cfg, ncol)
        if chunk is None:
            out = b.update_steps_host(P, E, names)
        else:
            out = b.update_steps_host_stream(P, E, names, chunk_steps=chunk, output_dtype=out_dtype)
        fr = b.get_fronts()
        b.finalize()
        return out, fr

    ref, fr_ref = run(Pm, Em)
    # the single call copies its forcing in chunks while the window runs (frontier of arrived hours): everything first, 7 h at a time
    for hours in ("0", "7", "529"):
        out, fr = run(Pm, Em, env={"LGARB_HOST_CHUNK_HOURS": hours})
        for n in names:
            assert np.array_equal(out[n], ref[n]), (hours, n)
        for k in ("depth", "theta", "psi", "nfronts"):
            assert np.array_equal(fr[k], fr_ref[k], equal_nan=True), (hours, k)
    for chunk in (0, 530, 200, 64):
        out, fr = run(Pm, Em, chunk)
        for n in names:
            assert np.array_equal(out[n], ref[n]), (chunk, n)
        for k in ("depth", "theta", "psi", "nfronts"):
            assert np.array_equal(fr[k], fr_ref[k], equal_nan=True), (chunk, k)
    P32, E32 = Pm.astype(np.float32), Em.astype(np.float32)
    ref32, _ = run(P32.astype(np.float64), E32.astype(np.float64))
    out32, _ = run(P32, E32, 150)
    for n in names:
        assert np.array_equal(out32[n], ref32[n]), n
    outf, _ = run(Pm, Em, 150, np.float32)
    for n in names:
        assert outf[n].dtype == np.float32 and np.array_equal(outf[n], ref[n].astype(np.float32)), n


def test_heterogeneous_batch_against_oracle(cuda_device, oracle, tmp_path):
    """BASELINE config 3 in small: Bushland layering, per-column soil triples drawn from the PARSED
    vG_params_stat_nom_ordered.dat (quirk Q2), per-column circular time shift of the forcing."""
    from oracle.oracle_py import DATA, write_config, read_forcing
    ncol, T = 192, 1200
    rng = np.random.default_rng(12345)
    types = rng.integers(1, 13, size=(ncol, 3)).astype(np.int32)
    shifts = (np.arange(ncol) * 7919) % 8760
    cfg = write_config(DATA / "configs" / "config_lasam_Bushland.txt", tmp_path / "c3.txt",
                       soil_params_file=str(DATA / "vG_params_stat_nom_ordered.dat"), max_valid_soil_types="12", layer_soil_type="1,2,3")
    P, E = read_forcing(DATA / "forcing" / "forcing_data_resampled_uniform_Bushland.csv")
    P, E = P[:8760], E[:8760]
    import torch
    import lgar_c_b200
    Pm = np.stack([np.roll(P, -int(s))[:T] for s in shifts], axis=1)
    Em = np.stack([np.roll(E, -int(s))[:T] for s in shifts], axis=1)
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol, 0, types)
    dP = torch.from_numpy(np.ascontiguousarray(Pm)).cuda(); dE = torch.from_numpy(np.ascontiguousarray(Em)).cuda()
    sinks = torch.zeros((12, T, ncol), dtype=torch.float64, device="cuda"); nfs = torch.zeros((T, ncol), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()  # the engine runs on its own stream: inputs/sinks made by torch must be complete
    b.update_steps_device(T, dP.data_ptr(), dE.data_ptr(), ncol, {n: sinks[k].data_ptr() for k, n in enumerate(lgar_c_b200.OUTPUT_SCALARS)}, ncol, nfs.data_ptr())
    b.synchronize()
    scal = sinks.cpu().numpy().transpose(1, 0, 2); nf = nfs.cpu().numpy()
    assert (b.get_status() & 0xFFFF == 0).all()
    fr = b.get_fronts()
    for c in range(ncol):
        p = oracle.params_from_config(cfg, layer_soil_type=[int(x) for x in types[c]])
        s = oracle.new_state(p)
        r = oracle.run_series(p, s, Pm[:, c], Em[:, c], cap=16)
        assert r["done"] == T
        assert np.array_equal(nf[:, c], r["nfronts"]), f"col {c} soils {types[c]}: front count differs at step {int(np.argmax(nf[:, c] != r['nfronts']))}"
        assert_scalars_close(scal[:, :, c], r["scalars"], f"col {c} soils {types[c]}")
        arr, flags, n = fronts_record(fr, c)
        assert_fronts_close(arr, flags, r["fronts"][-1], r["flags"][-1], n, f"col {c} final fronts")


def test_replicated_columns_are_identical_100k(cuda_device, tmp_path):
    """BASELINE config 2: a synthetic case replicated to 100 000 columns on one GPU -- every column
    must equal column 0 bit for bit, and column 0 must match the reference's golden vector."""
    g = load_golden("synth_1")
    cfg = config_for("synth_1", tmp_path)
    import torch
    import lgar_c_b200
    ncol, T = 100_000, len(g["precip"])
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol)
    dP = torch.from_numpy(np.repeat(g["precip"][:, None], ncol, axis=1).copy()).cuda()
    dE = torch.from_numpy(np.repeat(g["pet"][:, None], ncol, axis=1).copy()).cuda()
    torch.cuda.synchronize()  # the engine runs on its own stream: inputs/sinks made by torch must be complete
    b.update_steps_device(T, dP.data_ptr(), dE.data_ptr(), ncol)
    b.synchronize()
    for name in ("soil_storage", "infiltration", "surface_runoff", "mass_balance"):
        v = b.get_value(name)
        assert (v == v[0]).all(), name
    assert (b.get_value("soil_num_wetting_fronts") == g["nfronts"][-1]).all()
    assert b.get_value("soil_storage")[0] == g["scalars"][-1, 5]
    mb = b.get_global_mass_balance(0, 1000)
    assert np.abs(mb[:, 15]).max() < 1e-6     # global balance closes (reference: 1.9e-9 cm on this case)


@pytest.mark.parametrize("name", ["synth_1", "synth_2"])
def test_synthetic_cases_100k_against_hydrus_and_lgar_py(cuda_device, name, tmp_path):
    """BASELINE config 2 in full: synth_1 / synth_2 replicated to 100 000 columns, per-step soil_storage and surface_runoff of
    every column identical to column 0, column 0 within 1e-9 of the reference's golden vector and within the millimetre-level
    bars of the HYDRUS-1D and LGAR-Py solutions the reference ships (helpers.assert_physics_sanity)."""
    from helpers import assert_physics_sanity
    g = load_golden(name)
    cfg = config_for(name, tmp_path)
    import torch
    import lgar_c_b200
    ncol, T = 100_000, len(g["precip"])
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol)
    dP = torch.from_numpy(np.repeat(g["precip"][:, None], ncol, axis=1).copy()).cuda()
    dE = torch.from_numpy(np.repeat(g["pet"][:, None], ncol, axis=1).copy()).cuda()
    stor = torch.zeros((T, ncol), dtype=torch.float64, device="cuda")
    ro = torch.zeros((T, ncol), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()  # the engine runs on its own stream: inputs/sinks made by torch must be complete
    b.update_steps_device(T, dP.data_ptr(), dE.data_ptr(), ncol, {"soil_storage": stor.data_ptr(), "surface_runoff": ro.data_ptr()}, ncol)
    b.synchronize()
    s, r = stor.cpu().numpy(), ro.cpu().numpy()
    assert (s == s[:, :1]).all() and (r == r[:, :1]).all()
    for got, k in ((s[:, 0], 5), (r[:, 0], 3)):
        want = g["scalars"][:, k]
        assert np.array_equal(got, want), f"{name}: scalar {k} differs from the reference's golden vector"
    assert_physics_sanity(name, s[:, 0], r[:, 0])


def test_bmi_variable_access_semantics(cuda_device, tmp_path):
    import lgar_c_b200
    b = lgar_c_b200.BmiLGARBatch()
    ncol = 7
    b.initialize(config_for("Phillipsburg", tmp_path), ncol)
    assert b.get_grid_size(1) == ncol and b.get_grid_size(2) == ncol * 3 and b.get_var_nbytes("soil_storage") == 8 * ncol
    # forcing defaults to -1 until set (reference src/lgar.cxx:172-175)
    assert (b.get_value("precipitation_rate") == -1.0).all()
    # soil_depth_layers: cumulative thickness from index 0, in cm (reference quirk Q10, src/bmi_lgar.cxx:1350)
    assert np.array_equal(b.get_value("soil_depth_layers")[3], [0.0, 44.0, 175.0])
    assert list(b.get_value("soil_num_wetting_fronts")) == [3] * ncol
    b.set_value("precipitation_rate", np.zeros(ncol))
    b.set_value("potential_evapotranspiration_rate", np.full(ncol, 0.1))
    b.set_value_at_indices("precipitation_rate", [2, 5], [4.0, 12.0])
    assert list(b.get_value_at_indices("precipitation_rate", [5, 2, 0])) == [12.0, 4.0, 0.0]
    b.update_until(3600.0)
    with pytest.raises(lgar_c_b200.LgarbError):
        b.update_until(0.0)       # reference: assert (t > 0.0), src/bmi_lgar.cxx:901
    infil = b.get_value("infiltration")
    assert infil[0] == 0.0 and infil[2] > 0.0 and infil[5] > infil[2]
    assert list(b.get_value_range("infiltration", 2, 2)) == [infil[2], infil[3]]
    with pytest.raises(lgar_c_b200.LgarbError, match="does not exist"):
        b.get_value("nonexistent_variable")      # reference: throw, src/bmi_lgar.cxx:1421-1425
    import torch
    ptr = b.get_value_ptr("infiltration")
    assert ptr != 0
    out = torch.zeros(ncol, dtype=torch.float64, device="cuda")
    b.get_value_device("infiltration", out.data_ptr())
    assert np.array_equal(out.cpu().numpy(), infil)
    # negative forcing is clamped with a warning bit (reference src/bmi_lgar.cxx:325-337), never fatal
    b.set_value("precipitation_rate", np.full(ncol, -2.0))
    b.update()
    st = b.get_status()
    assert (st & lgar_c_b200.STATUS_FAIL_MASK == 0).all() and (st & 0x10000).all()
    b.finalize()
    with pytest.raises(lgar_c_b200.LgarbError):
        b.update()


def test_invalid_soil_columns_pass_precip_through(cuda_device, tmp_path):
    """Columns whose soil type exceeds max_valid_soil_types return Q = precip (src/bmi_lgar.cxx:145-179)
    while their neighbours run normally."""
    import lgar_c_b200
    b = lgar_c_b200.BmiLGARBatch()
    types = np.array([[13, 14, 15], [13, 26, 15], [13, 14, 15]], dtype=np.int32)
    b.initialize(config_for("Phillipsburg", tmp_path), 3, 0, types)
    b.set_value("precipitation_rate", [3.0, 3.0, 3.0])
    b.set_value("potential_evapotranspiration_rate", [0.2, 0.2, 0.2])
    b.update()
    q = b.get_value("total_discharge"); r = b.get_value("surface_runoff"); i = b.get_value("infiltration")
    assert q[1] == r[1] == 3.0 * 0.001 and i[1] == 0.0
    assert i[0] == i[2] and i[0] > 0.0
    assert list(b.get_value("soil_num_wetting_fronts")) == [4, 0, 4]


def test_hourly_update_through_stage_kernels(cuda_device, tmp_path):
    """The BMI single-step verb on a large batch (LGARB_UPDATE_WAVE_MIN_COLS: the stream kernel finishes the cached-flux columns
    and lists the others, the listed columns take the stage kernels as a window of one hour, slots drawing columns from the
    list) against the stream/active kernel pair: SetValue x2 -> Update -> GetValue per hour, every output of every hour, the
    final fronts and the global balance bit-identical -- also with fewer slots than listed columns (slots take a second column)."""
    import os
    import bench
    import lgar_c_b200
    ncol, T = 20_000, 120
    cfg = bench.write_workload_config(tmp_path)
    soils, layers = bench.column_soils(0, ncol), bench.column_layers(0, ncol)
    P, E = bench.host_forcing(ncol, 0, T, 2026)
    P[30:34] = np.maximum(P[30:34], 6.0)     # a storm on every column: every column listed at once
    res = []
    for env in ({}, {"LGARB_UPDATE_WAVE_MIN_COLS": "0"}, {"LGARB_UPDATE_WAVE_MIN_COLS": "0", "LGARB_UPDATE_WAVE_SLOTS": "4096"}):
        os.environ.update(env)
        try:
            b = lgar_c_b200.BmiLGARBatch()
        finally:
            for k in env:
                del os.environ[k]
        b.initialize(cfg, ncol, 0, soils, layers)
        out = np.empty((T, len(lgar_c_b200.OUTPUT_SCALARS), ncol))
        nf = np.empty((T, ncol), dtype=np.int32)
        for t in range(T):
            b.set_value("precipitation_rate", P[t])
            b.set_value("potential_evapotranspiration_rate", E[t])
            b.update()
            for k, name in enumerate(lgar_c_b200.OUTPUT_SCALARS):
                out[t, k] = b.get_value(name)
            nf[t] = b.get_value("soil_num_wetting_fronts")
        res.append((out, nf, b.get_status(), b.get_fronts(), b.get_global_mass_balance(), b.get_current_time(), b.get_launch_count()))
        del b
    o0, n0, s0, f0, m0, t0, l0 = res[0]
    assert (s0 & lgar_c_b200.STATUS_FAIL_MASK == 0).all()
    for (o1, n1, s1, f1, m1, t1, l1) in res[1:]:
        assert l1 > l0, "the stage kernels did not run"
        assert t1 == t0 and np.array_equal(s0, s1) and np.array_equal(n0, n1)
        assert np.array_equal(o0, o1, equal_nan=True) and np.array_equal(m0, m1, equal_nan=True)
        for k in ("depth", "theta", "psi", "K", "dzdt", "layer", "to_bottom", "nfronts"):
            assert np.array_equal(f0[k], f1[k], equal_nan=True), k


def test_engine_on_a_block_of_global_host_arrays(cuda_device, tmp_path):
    """lgarb_update_steps_host_pitched: an engine that owns columns [col0, col0 + n) of global [step][n_total] host arrays (what
    one device of a multi-device group does, include/lgarb_group.h) gives the outputs of an engine fed that block alone, bit for
    bit, writes nothing outside its block, and its balance terms computed on the device equal the host version."""
    import torch
    import lgar_c_b200
    g = load_golden("Phillipsburg_with_reservoir")
    cfg = config_for("Phillipsburg_with_reservoir", tmp_path)
    ntot, col0, ncol, T = 333, 77, 130, 300
    shifts = (np.arange(ntot) * 7919) % 8760
    Pm = np.ascontiguousarray(np.stack([np.roll(g["precip"], -int(s))[:T] for s in shifts], axis=1))
    Em = np.ascontiguousarray(np.stack([np.roll(g["pet"], -int(s))[:T] for s in shifts], axis=1))
    names = ["total_discharge", "soil_storage"]
    a = lgar_c_b200.BmiLGARBatch()
    a.initialize(cfg, ncol)
    ref = a.update_steps_host(Pm[:, col0:col0 + ncol], Em[:, col0:col0 + ncol], names)
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol)
    out = b.update_steps_host_block(Pm, Em, col0, names)
    for n in names:
        assert np.array_equal(out[n][:, col0:col0 + ncol], ref[n]), n
        assert np.isnan(out[n][:, :col0]).all() and np.isnan(out[n][:, col0 + ncol:]).all(), "wrote outside its block"
    d_out = torch.zeros((ncol, 16), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    b._ck(b._L.lgarb_get_global_mass_balance_device(b._h, d_out.data_ptr()))
    b.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), b.get_global_mass_balance(), equal_nan=True)


def test_group_of_one_device(cuda_device, tmp_path):
    """liblgarb_group.so with a single device (no NCCL traffic): the group verbs give what the engine gives."""
    import lgar_c_b200
    from lgar_c_b200.group import LgarGroup
    g = load_golden("Phillipsburg_with_reservoir")
    cfg = config_for("Phillipsburg_with_reservoir", tmp_path)
    ncol, T = 200, 250
    shifts = (np.arange(ncol) * 7919) % 8760
    Pm = np.ascontiguousarray(np.stack([np.roll(g["precip"], -int(s))[:T] for s in shifts], axis=1))
    Em = np.ascontiguousarray(np.stack([np.roll(g["pet"], -int(s))[:T] for s in shifts], axis=1))
    a = lgar_c_b200.BmiLGARBatch()
    a.initialize(cfg, ncol)
    ref = a.update_steps_host(Pm, Em, ["total_discharge"])
    grp = LgarGroup([0])
    grp.initialize(cfg, ncol)
    assert grp.shards() == [(0, ncol)]
    out = grp.update_steps_host(Pm, Em, ["total_discharge"])
    assert np.array_equal(out["total_discharge"], ref["total_discharge"])
    gm, tot, ms = grp.gather_mass_balance()
    want = a.get_global_mass_balance()
    assert np.array_equal(gm, want, equal_nan=True)
    assert np.allclose(tot, want.sum(axis=0), rtol=1e-12, atol=1e-9)
    grp.finalize()
