"""lgar_c_b200/csrc/lgar_pow.cuh against the host C library's pow, bit for bit.

The reference calls libm pow() for every van Genuchten relation (reference src/soil_funcs.cxx:147-187); the engine's pow restates
that function's algorithm (glibc >= 2.28, FMA build) so that the device returns the same double.  CPU half: the header compiled
by the host compiler (LGAR_HOSTSIM) against pow() of the libm the oracle and oracle/_ref link.  GPU half (test_gpu_parity.py::
test_device_pow_equals_host_pow) sends the same kind of samples through lgarb_debug_pow."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = r'''
#define LGAR_HOSTSIM
#include "lgar_pow.cuh"
extern "C" void hs_pow(const double* x, const double* y, double* out, long n) { for (long i = 0; i < n; i++) out[i] = lgp_pow(x[i], y[i]); }
extern "C" void hs_try_pow(const double* x, const double* y, double* out, int* ok, long n) {
  for (long i = 0; i < n; i++) { double r = 0; ok[i] = lgp_try_pow(x[i], y[i], &r); out[i] = r; }
}
extern "C" void libm_pow(const double* x, const double* y, double* out, long n) { for (long i = 0; i < n; i++) { volatile double a = x[i], b = y[i]; out[i] = pow(a, b); } }
'''


def path_samples(n, seed):
    """(x, y) of the shapes the path produces: (alpha h)^n, (1+t)^m, Se^(1/m), Se^(-1/m), (.)^(1/n), S^b, x^2, x^3, Se^(3+1/lambda)"""
    r = np.random.default_rng(seed)
    u = lambda: r.random(n)
    xs = [np.exp(-12 + 30 * u()), 1 + np.exp(-30 + 60 * u()), u(), u(), np.exp(-40 + 80 * u()), 100 * u(),
          np.exp(-20 + 40 * u()) * np.where(u() < 0.01, -1, 1), np.exp(-700 + 1400 * u())]
    ys = [1.01 + 3 * u(), 0.01 + 0.9 * u(), 1 + 50 * u(), -(1 + 50 * u()), 0.2 + 0.8 * u(), np.full(n, 2.5),
          np.where(u() < 0.5, 2.0, 3.0), -40 + 80 * u()]
    return np.concatenate(xs), np.concatenate(ys)


def special_samples():
    sp = np.array([0.0, -0.0, 1.0, -1.0, 2.0, -2.0, 0.5, -0.5, 3.0, -3.0, 1e-320, -1e-320, 5e-324, 1e308, -1e308, np.inf, -np.inf, np.nan,
                   1e-70, -1e-70, 1e19, -1e19, 9007199254740993.0, 4503599627370497.0, -4503599627370497.0, 1.5, -1.5, 2.5, 2.0**-1022,
                   np.finfo(np.float64).max, 1075.0, -1075.0, 1e-300, 1e300, 0.9999999999999999, 1.0000000000000002, 2.0**63, 2.0**-65, 2.0**-66,
                   -2.0**63])
    x, y = np.meshgrid(sp, sp, indexing="ij")
    return x.ravel().copy(), y.ravel().copy()


def edge_samples(n, seed):
    """results in the subnormal range and next to overflow; random bit patterns"""
    r = np.random.default_rng(seed)
    x = np.exp(-2 + 4 * r.random(n))
    t = np.where(r.random(n) < 0.5, -1, 1) * (700 + 50 * r.random(n))
    y = t / np.log(x)
    xb = r.integers(0, 2**64, n, dtype=np.uint64).view(np.float64)
    yb = r.integers(0, 2**64, n, dtype=np.uint64).view(np.float64)
    return np.concatenate([x, xb]), np.concatenate([y, yb])


def same_bits(a, b):
    nan = np.isnan(a) & np.isnan(b)
    return (a.view(np.uint64) == b.view(np.uint64)) | nan


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("pow")
    (d / "p.cpp").write_text(SRC)
    so = d / "libp.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", str(ROOT / "lgar_c_b200" / "csrc"), "-o", str(so), str(d / "p.cpp")])
    return C.CDLL(str(so))


def call(fn, x, y):
    out = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    fn(x.ctypes.data_as(dp), y.ctypes.data_as(dp), out.ctypes.data_as(dp), C.c_long(len(x)))
    return out


@pytest.mark.parametrize("maker", [lambda: path_samples(1_000_000, 1), special_samples, lambda: edge_samples(1_000_000, 2)])
def test_restated_pow_equals_libm_pow(lib, maker):
    x, y = maker()
    a, b = call(lib.libm_pow, x, y), call(lib.hs_pow, x, y)
    bad = ~same_bits(a, b)
    assert not bad.any(), [(float(x[i]).hex(), float(y[i]).hex(), float(a[i]).hex(), float(b[i]).hex()) for i in np.flatnonzero(bad)[:5]]


def test_inlined_common_case_equals_full_function(lib):
    x, y = path_samples(500_000, 3)
    out = np.empty_like(x)
    ok = np.zeros(len(x), dtype=np.int32)
    dp = C.POINTER(C.c_double)
    lib.hs_try_pow(x.ctypes.data_as(dp), y.ctypes.data_as(dp), out.ctypes.data_as(dp), ok.ctypes.data_as(C.POINTER(C.c_int)), C.c_long(len(x)))
    full = call(lib.hs_pow, x, y)
    m = ok != 0
    assert m.mean() > 0.8                       # it is the common case
    assert same_bits(out[m], full[m]).all()


def test_tables_are_those_of_this_libm(tmp_path):
    """The committed tables are the ones tools/extract_libm_pow_tables.py finds in the libm of this image."""
    import sys
    out = tmp_path / "t.h"
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "extract_libm_pow_tables.py"), str(out)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("this libm does not carry the table-driven pow: " + (r.stdout + r.stderr)[-200:])
    assert out.read_text() == (ROOT / "lgar_c_b200" / "csrc" / "lgar_pow_tables.h").read_text()
