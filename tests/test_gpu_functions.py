"""Frozen-snapshot kernel tests, one per function group: the leaf functions of the path evaluated
element by element on the device (lgarb_debug_functions) against the oracle's exported leaf functions (oracle/lgar_oracle.h,
"exported for element-wise kernel tests") with the same inputs -- before any control flow of the step is involved.  The
snapshot is a heterogeneous batch of the bench workload (random soil triples, two layerings) after 300 forcing hours, so the
list functions see developed front lists.  Bit equality: the device pow returns the host library's double (lgar_pow.cuh).
Needs a B200."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NCOL, T = 2048, 300


@pytest.fixture(scope="module")
def snapshot(cuda_device, oracle, tmp_path_factory):
    import torch
    import bench
    import lgar_c_b200
    tmp = tmp_path_factory.mktemp("functions")
    cfg = bench.write_workload_config(tmp)
    soils, layers = bench.column_soils(0, NCOL), bench.column_layers(0, NCOL)
    P, E = bench.host_forcing(NCOL, 0, T, 31337)
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, NCOL, 0, soils, layers)
    dP, dE = torch.from_numpy(P).cuda(), torch.from_numpy(E).cuda()
    torch.cuda.synchronize()
    b.update_steps_device(T, dP.data_ptr(), dE.data_ptr(), NCOL)
    b.synchronize()
    assert (b.get_status() & lgar_c_b200.STATUS_FAIL_MASK == 0).all()
    params = [oracle.params_from_config(cfg, layer_soil_type=[int(x) for x in soils[c]], layer_thickness=[float(x) for x in layers[c]])
              for c in range(NCOL)]
    return dict(b=b, params=params, fronts=b.get_fronts(), oracle=oracle)


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)


def _assert_bit_equal(got, want, what):
    bad = np.nonzero(_bits(got) != _bits(want))[0]
    assert bad.size == 0, f"{what}: {bad.size} of {got.size} items differ, first at {bad[0]}: {got[bad[0]]!r} vs {want[bad[0]]!r}"


@pytest.mark.parametrize("layer", [1, 2, 3])
def test_van_genuchten_functions(snapshot, layer):
    """calc_theta_from_h, calc_Se_from_h, calc_K_from_Se, calc_h_from_Se (reference src/soil_funcs.cxx:147-179) for every column's
    soil of `layer`, over heads from 1e-3 to 1e7 cm and saturations over (0, 1]."""
    b, params, L = snapshot["b"], snapshot["params"], snapshot["oracle"].lib
    rng = np.random.default_rng(100 + layer)
    h = 10.0 ** rng.uniform(-3, 7, NCOL)
    h[:8] = [0.0, 1e-300, 1.0, 15495.0, 340.9, 1e7, 1e7 * (1 + 2e-16), 3.3e-9]
    Se = rng.uniform(1e-9, 1.0, NCOL)
    Se[:6] = [1.0, 1.0 - 2 ** -53, 0.5, 1e-12, 0.999999999, 0.05]
    soil = [p.soil[min(layer, p.num_layers)] for p in params]
    _assert_bit_equal(b.debug_functions(0, h, layer=layer), np.array([L.lgo_theta_from_h(x, C.byref(s)) for x, s in zip(h, soil)]), "theta_from_h")
    _assert_bit_equal(b.debug_functions(1, h, layer=layer), np.array([L.lgo_Se_from_h(x, s.alpha, s.m, s.n) for x, s in zip(h, soil)]), "Se_from_h")
    _assert_bit_equal(b.debug_functions(2, Se, layer=layer), np.array([L.lgo_K_from_Se(x, s.Ksat, s.m) for x, s in zip(Se, soil)]), "K_from_Se")
    _assert_bit_equal(b.debug_functions(3, Se, layer=layer), np.array([L.lgo_h_from_Se(x, s.alpha, s.m, s.n) for x, s in zip(Se, soil)]), "h_from_Se")


def test_geff_closed_form(snapshot):
    """calc_Geff (reference src/soil_funcs.cxx:15-131), the form the configuration selects, for pairs theta1 < theta2 inside each
    soil's range and the degenerate pairs the front logic produces (equal, reversed, at the limits)."""
    b, params, L = snapshot["b"], snapshot["params"], snapshot["oracle"].lib
    rng = np.random.default_rng(7)
    soil = [p.soil[1] for p in params]
    lo = np.array([s.theta_r for s in soil])
    hi = np.array([s.theta_e for s in soil])
    u = np.sort(rng.uniform(0.0, 1.0, (2, NCOL)), axis=0)
    t1 = lo + u[0] * (hi - lo)
    t2 = lo + u[1] * (hi - lo)
    t1[:4], t2[:4] = [lo[0], t2[1], hi[2], t2[3]], [hi[0], t2[1], lo[2], hi[3]]
    want = np.array([L.lgo_Geff(p.use_closed_form_G, x, y, C.byref(s), p.nint) for x, y, s, p in zip(t1, t2, soil, params)])
    _assert_bit_equal(b.debug_functions(4, t1, t2, layer=1), want, "Geff")


def test_list_functions_on_the_snapshot(snapshot):
    """lgar_calc_mass_bal (reference src/lgar.cxx:2632-2668) and calc_aet (src/aet.cxx:17-77) on every column's front list as it
    stands after 300 hours."""
    from oracle.oracle_py import Front
    b, params, fr, L = snapshot["b"], snapshot["params"], snapshot["fronts"], snapshot["oracle"].lib
    rng = np.random.default_rng(11)
    pet = rng.uniform(0.0, 0.08, NCOL)
    dt = rng.choice([1.0, 1.0 / 12.0], NCOL)
    want_mb, want_aet = np.empty(NCOL), np.empty(NCOL)
    assert int(fr["nfronts"].max()) >= 5, "the snapshot should hold developed front lists"
    for c, p in enumerate(params):
        n = int(fr["nfronts"][c])
        arr = (Front * n)()
        for i in range(n):
            arr[i].depth, arr[i].theta, arr[i].psi = fr["depth"][c][i], fr["theta"][c][i], fr["psi"][c][i]
            arr[i].K, arr[i].dzdt = fr["K"][c][i], fr["dzdt"][c][i]
            arr[i].layer, arr[i].to_bottom = int(fr["layer"][c][i]), int(fr["to_bottom"][c][i])
        want_mb[c] = L.lgo_calc_mass_bal(p.cum_thickness, arr, n)
        want_aet[c] = L.lgo_calc_aet(pet[c], dt[c], p.wilting_point_psi_cm, p.field_capacity_psi_cm, arr, C.byref(p.soil[int(arr[0].layer)]))  # the soil of the head front's layer
    _assert_bit_equal(b.debug_functions(5, pet), want_mb, "lgar_calc_mass_bal")
    _assert_bit_equal(b.debug_functions(6, pet, dt), want_aet, "calc_aet")


def test_conceptual_reservoir(snapshot):
    """calc_CR_Q (reference src/conceptual_reservoir.cxx:7-45) from every column's reservoir storages after 300 hours: discharge and
    both storages after the call."""
    b, params, L = snapshot["b"], snapshot["params"], snapshot["oracle"].lib
    rng = np.random.default_rng(13)
    dt = rng.choice([1.0, 1.0 / 12.0], NCOL)
    inflow = np.where(rng.random(NCOL) < 0.5, 0.0, rng.uniform(0.0, 3.0, NCOL))
    fast0, slow0 = b.debug_functions(10, dt), b.debug_functions(11, dt)   # the storages as they stand in the state (cm)
    assert (fast0 > 0.01).any() and (fast0 < 0.01).any(), "both branches of the fast bucket should be reached"
    want = np.empty((3, NCOL))
    for c, p in enumerate(params):
        f, s = C.c_double(fast0[c]), C.c_double(slow0[c])
        want[0, c] = L.lgo_calc_CR_Q(dt[c], p.a, p.a_slow, p.b, p.b_slow, p.frac_slow, inflow[c], C.byref(f), C.byref(s))
        want[1, c], want[2, c] = f.value, s.value
    got = np.stack([b.debug_functions(k, dt, inflow) for k in (7, 8, 9)])
    for k, name in enumerate(("Q", "fast storage", "slow storage")):
        _assert_bit_equal(got[k], want[k], f"calc_CR_Q {name}")
