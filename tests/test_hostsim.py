"""The kernel TEXT (lgar_c_b200/csrc/lgar_physics.cuh + lgar_step.cuh) compiled for the host by
tests/hostsim and run against the oracle: a CPU-side check of the device code's logic (slot
arithmetic, the two-kernel split, the epilogue).  Test infrastructure only -- the product library
contains no host path.  With lgar_pow.cuh underneath (the default build: the host library's pow restated operation for operation)
the kernel text reproduces the reference bit for bit; nothing is left to a tolerance."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import STANDARD, RANDOM, load_golden, config_for, assert_scalars_close, assert_fronts_close

HS = Path(__file__).resolve().parent / "hostsim"
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def hostsim():
    so = HS / "libhostsim.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", str(so), str(HS / "hostsim.cpp")])
    lib = C.CDLL(str(so))
    lib.hostsim_run_series.argtypes = [C.c_char_p, _ip, _dp, C.c_int, _dp, _dp, _dp, _ip, C.c_int, _dp, _ip, C.c_int, _ip, C.c_char_p, C.c_int]
    lib.hostsim_run_series.restype = C.c_int
    return lib


def run(lib, cfg, P, E, force, cap=16):
    P = np.ascontiguousarray(P, dtype=np.float64)
    E = np.ascontiguousarray(E, dtype=np.float64)
    n = len(P)
    scal = np.zeros((n, 12)); nf = np.zeros(n, dtype=np.int32); fr = np.zeros((n, cap, 5)); fl = np.zeros((n, cap), dtype=np.int32)
    st = C.c_int(0)
    err = C.create_string_buffer(512)
    done = lib.hostsim_run_series(str(cfg).encode(), None, None, n, P.ctypes.data_as(_dp), E.ctypes.data_as(_dp), scal.ctypes.data_as(_dp),
                                  nf.ctypes.data_as(_ip), cap, fr.ctypes.data_as(_dp), fl.ctypes.data_as(_ip), force, C.byref(st), err, 512)
    assert done >= 0, err.value.decode()
    return done, scal, nf, fr, fl, st.value


@pytest.mark.parametrize("mode", [2, 3, 4, 5])
@pytest.mark.parametrize("name", STANDARD + RANDOM)
def test_window_stepping_equals_hourly_stepping(hostsim, name, mode, tmp_path):
    """The window kernel's per-column functions (modes 2/3: advance through cached hours in registers,
    active step at the column's own hour) and the staged lane program of lgar_lanes.cuh (modes 4/5:
    FETCH/HEAD/FRONT/SEARCH/TAIL) must reproduce hour-by-hour stepping bit for bit."""
    g = load_golden(name)
    n = min(len(g["precip"]), 2000)
    cfg = config_for(name, tmp_path)
    d0, s0, nf0, fr0, fl0, st0 = run(hostsim, cfg, g["precip"][:n], g["pet"][:n], 0)
    d1, s1, nf1, fr1, fl1, st1 = run(hostsim, cfg, g["precip"][:n], g["pet"][:n], mode)
    assert d0 == d1 == n and st0 == st1
    assert np.array_equal(nf0, nf1) and np.array_equal(s0, s1)
    assert np.array_equal(fr0[n - 1], fr1[n - 1]) and np.array_equal(fl0[n - 1], fl1[n - 1])


@pytest.mark.parametrize("force", [0, 1])
@pytest.mark.parametrize("name", STANDARD + RANDOM)
def test_kernel_text_matches_golden(hostsim, name, force, tmp_path):
    g = load_golden(name)
    n = min(len(g["precip"]), 3000)
    done, scal, nf, fr, fl, status = run(hostsim, config_for(name, tmp_path), g["precip"][:n], g["pet"][:n], force)
    assert done == n and status & 0xFFFF == 0
    assert np.array_equal(nf, g["nfronts"][:n]), "front counts differ from the reference"
    # the default build's pow IS the reference's (lgar_pow.cuh): no tolerance
    assert np.array_equal(scal, g["scalars"][:n]), f"{name}: {np.count_nonzero(scal != g['scalars'][:n])} scalars differ in the last bits"
    for i, t in enumerate(g["ck_steps"]):
        if t < n:
            k = int(nf[t])
            assert np.array_equal(fl[t][:k], g["ck_flags"][i][:k])
            assert np.array_equal(fr[t][:k], g["ck_fronts"][i][:k]), f"{name}@{t}: front fields differ in the last bits"


def test_fast_search_path_reproduces_the_plain_walk(tmp_path):
    """lgar_search.cuh: every psi search and every correction search of every golden case through BOTH the fast path (root by
    Newton, bracket, replay of the digit walk with the evaluations outside the bracket skipped) and the plain trip-by-trip walk
    of the reference (-DLGAR_FF_VERIFY): the final records -- psi, theta, mass error, trip count, the fronts a correction search
    rewrites -- must be equal bit for bit, and the fast path must be what actually runs (>= 95 % of the searches finished by it)."""
    so = tmp_path / "libhostsim_ffv.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-DLGAR_FF_STATS", "-DLGAR_FF_VERIFY", "-o", str(so),
                           str(HS / "hostsim.cpp")])
    lib = C.CDLL(str(so))
    lib.hostsim_run_series.argtypes = [C.c_char_p, _ip, _dp, C.c_int, _dp, _dp, _dp, _ip, C.c_int, _dp, _ip, C.c_int, _ip, C.c_char_p, C.c_int]
    lib.hostsim_run_series.restype = C.c_int
    for name in STANDARD + RANDOM:
        g = load_golden(name)
        n = min(len(g["precip"]), 3000)
        for force in (0, 4):
            done, scal, nf, fr, fl, status = run(lib, config_for(name, tmp_path), g["precip"][:n], g["pet"][:n], force)
            assert done == n and np.array_equal(nf, g["nfronts"][:n]) and np.array_equal(scal, g["scalars"][:n]), (name, force)
    st = (C.c_longlong * 8)()
    lib.hostsim_ff_stats(st, 0)
    searches, ff_done, declined, fallback, evals_ff, _, trips, verify_bad = list(st)
    assert verify_bad == 0, f"{verify_bad} searches ended differently through the fast path"
    assert searches > 10_000 and ff_done >= 0.95 * searches, (searches, ff_done, declined, fallback)
    assert evals_ff < 0.35 * trips, "the fast path evaluates more than a third of the trips it replays"


@pytest.mark.parametrize("force,ff", [(0, 1), (0, 0), (4, 1), (4, 0)])
def test_path_counters_equal_the_oracles(hostsim, oracle, force, ff, tmp_path):
    """The kernel text's path counters (LGAR_PATH_* of lgar_types.h: merge, layer crossing, domain-bottom crossing, dry-over-wet,
    theta < theta_r deletion, saturated-front depth loop, surficial front, every exit of the psi search, correction searches,
    cached / active steps) against the oracle's (LGO_PATH_*), summed over all golden cases: equal branch by branch, in both
    program forms, with the psi searches through the fast path and through the plain walk."""
    hostsim.hostsim_set_use_ff(ff)
    buf = (C.c_ulonglong * 24)()
    hostsim.hostsim_get_paths(buf, 1)
    want = np.zeros(24, dtype=np.int64)
    for name in STANDARD + RANDOM:
        g = load_golden(name)
        n = min(len(g["precip"]), 3000)
        cfg = config_for(name, tmp_path)
        done, scal, nf, fr, fl, st = run(hostsim, cfg, g["precip"][:n], g["pet"][:n], force)
        assert done == n and np.array_equal(scal, g["scalars"][:n])
        p = oracle.params_from_config(cfg)
        s = oracle.new_state(p)
        oracle.run_series(p, s, np.ascontiguousarray(g["precip"][:n]), np.ascontiguousarray(g["pet"][:n]), want_fronts=False)
        want += np.array(s.n_path[:])
    hostsim.hostsim_get_paths(buf, 1)
    hostsim.hostsim_set_use_ff(1)
    got = np.array(buf[:], dtype=np.int64)
    assert np.array_equal(got, want), (got, want)
    for i in (0, 1, 2, 3, 5, 7, 9, 10, 11, 16, 17, 18):   # what the golden cases reach (the others: tests/test_gpu_scale_parity.py)
        assert got[i] > 0, i


@pytest.fixture(scope="module")
def hostsim_libm_pow():
    """The same kernel text with x^y = the library pow (-DLGAR_ACCURATE_POW; on the host that is glibc's pow, the function
    the reference calls) and pow(x,2)/pow(x,3) through it as well."""
    so = HS / "libhostsim_acc.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-DLGAR_ACCURATE_POW", "-o", str(so),
                           str(HS / "hostsim.cpp")])
    lib = C.CDLL(str(so))
    lib.hostsim_run_series.argtypes = [C.c_char_p, _ip, _dp, C.c_int, _dp, _dp, _dp, _ip, C.c_int, _dp, _ip, C.c_int, _ip, C.c_char_p, C.c_int]
    lib.hostsim_run_series.restype = C.c_int
    return lib


@pytest.mark.parametrize("force", [0, 4])
@pytest.mark.parametrize("name", STANDARD + RANDOM)
def test_kernel_text_with_library_pow_is_bit_exact(hostsim_libm_pow, name, force, tmp_path):
    """With the reference's own pow underneath, the kernel text (hour-by-hour functions, force 0; staged lane program,
    force 4) reproduces the unmodified reference BIT FOR BIT on every golden case: all 12 per-step scalars, front counts,
    and depth/theta/psi/K/dzdt/layer/to_bottom of every front at the checkpoints.  So the only thing between the engine
    and the reference is the evaluation of x^y (exp(y log x) on the device, <= 1e-14 relative per call) --
    the logic, the order of the arithmetic and every quirk of SURVEY.md 8c are identical."""
    g = load_golden(name)
    n = min(len(g["precip"]), 3000)
    done, scal, nf, fr, fl, status = run(hostsim_libm_pow, config_for(name, tmp_path), g["precip"][:n], g["pet"][:n], force)
    assert done == n and status & 0xFFFF == 0
    assert np.array_equal(nf, g["nfronts"][:n])
    assert np.array_equal(scal, g["scalars"][:n]), f"{name}: {np.count_nonzero(scal != g['scalars'][:n])} scalars differ in the last bits"
    for i, t in enumerate(g["ck_steps"]):
        if t < n and (force == 0 or t == n - 1):   # the window modes export the fronts at the end of the window only
            k = int(nf[t])
            assert np.array_equal(fl[t][:k], g["ck_flags"][i][:k])
            assert np.array_equal(fr[t][:k], g["ck_fronts"][i][:k]), f"{name}@{t}: front fields differ in the last bits"


def test_more_than_16_fronts(hostsim, hostsim_libm_pow, tmp_path):
    """Column 5588097 of BASELINE config 5 reaches 19 wetting fronts at hour 5641 of the year (tools/parity_sampling_full.py):
    the kernel text with its 32-front capacity and two flag words must follow the oracle through it."""
    import sys
    import torch
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import parity_sampling_full as pf
    from oracle.oracle_py import Oracle, write_config
    c = pf.Config(5, tmp_path)
    ids = np.array([5588097], dtype=np.int64)
    T = 5800
    P, E = (x.numpy()[:, 0] for x in c.forcing(torch.from_numpy(ids), 0, T))
    cfg = write_config(c.cfg, tmp_path / "deep.txt", layer_soil_type=",".join(map(str, c.soils(ids)[0])),
                       layer_thickness=",".join(map(str, c.layers(ids)[0])) + "[cm]")
    orc = Oracle()
    p = orc.params_from_config(cfg)
    s = orc.new_state(p)
    o = orc.run_series(p, s, P, E, cap=32)
    for force in (0, 4):
        done, scal, nf, fr, fl, status = run(hostsim, cfg, P, E, force, cap=32)
        assert done == T == o["done"] and status & 0xFFFF == 0
        assert nf.max() == 19 and np.array_equal(nf, o["nfronts"])
        assert np.array_equal(scal, o["scalars"]), "config5 column 5588097"
        t = T - 1 if force else int(np.argmax(nf))
        assert np.array_equal(fl[t][:int(nf[t])], o["flags"][t][:int(nf[t])])
        assert np.array_equal(fr[t][:int(nf[t])], o["fronts"][t][:int(nf[t])])
        # and bit for bit with the library pow underneath (the oracle calls the same glibc pow as the reference)
        done, scal, nf, fr, fl, status = run(hostsim_libm_pow, cfg, P, E, force, cap=32)
        assert done == T and np.array_equal(nf, o["nfronts"]) and np.array_equal(scal, o["scalars"])
        assert np.array_equal(fr[t][:int(nf[t])], o["fronts"][t][:int(nf[t])])


@pytest.mark.parametrize("args", [["--config", "5", "--columns", "48", "--hours", "400"],
                                  ["--config", "4", "--columns", "16", "--hours", "300"],
                                  ["--config", "3", "--columns", "8", "--hours", "120", "--stress"],
                                  ["--config", "5", "--columns", "24", "--hours", "400", "--reference"]])
def test_bitexact_sampling_tool(args):
    """tools/bitexact_sample.py at a small size: random columns of BASELINE configs 3/4/5 (and the stress variant: 300 s
    sub-steps, no flux caching, numeric Geff) -- the kernel text with the library pow equals the oracle, and with --reference
    the UNMODIFIED reference itself, bit for bit in both program forms.  The full-size runs (10 000 columns x the full duration per config) are in profiles/r1_bitexact_sample_*.json."""
    import json
    import sys
    tool = Path(__file__).resolve().parent.parent / "tools" / "bitexact_sample.py"
    if "--reference" in args and not (tool.parent.parent / "oracle" / "_ref" / "liblgar_ref.so").exists():
        pytest.skip("oracle/_ref (the unmodified reference, built where /root/reference exists) is not here")
    r = subprocess.run([sys.executable, str(tool), "--cores", "2"] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.loads(r.stdout[r.stdout.index("{"):])
    assert rep["bit_equal_columns_hourly"] == rep["columns"] == rep["bit_equal_columns_lanes"] and not rep["not_bit_equal"]


def test_kernel_text_under_sanitizers(tmp_path):
    """The reference's CI builds with -fsanitize=address,undefined (.github/workflows/build_and_run_tests.yml:35-56).  The same
    for the kernel text: physics + staged lane program compiled for the host with both sanitizers and run through the golden
    cases in all six program forms and through the 19-front column (fixed-capacity front arrays, nibble-packed flag words, the
    live-prefix bookkeeping: the places where an index error would hide on a GPU), and malformed configuration files through the
    product's host-side reader (lgar_config.hpp): a clean error that names the key, never a crash."""
    import os
    import sys
    rt = []
    for name in ("libasan.so", "libubsan.so"):
        path = subprocess.run(["gcc", f"-print-file-name={name}"], capture_output=True, text=True).stdout.strip()
        if not os.path.isabs(path):
            pytest.skip(f"{name} is not installed")
        rt.append(path)
    so = tmp_path / "libhostsim_san.so"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-DLGAR_ACCURATE_POW",
                           "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-o", str(so), str(HS / "hostsim.cpp")])
    env = dict(os.environ, LD_PRELOAD=":".join(rt), ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    r = subprocess.run([sys.executable, str(HS / "sanitizer_run.py"), str(so)], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "sanitizers clean" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
    # the checker itself: the plain-C oracle built the same way, every golden case bit for bit and the 19-front column
    oso = tmp_path / "liblgar_oracle_san.so"
    root = HS.parent.parent
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-g", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-fsanitize=address,undefined",
                           "-fno-sanitize-recover=undefined", "-o", str(oso), str(root / "oracle" / "lgar_oracle.c"), "-lm"])
    r = subprocess.run([sys.executable, str(HS / "sanitizer_run_oracle.py")], env=dict(env, LGAR_ORACLE_SO=str(oso)), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "oracle: sanitizers clean" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
