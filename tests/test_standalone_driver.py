"""lgar_c_b200/lasam_b200_standalone (the batched counterpart of the reference's src/bmi_main_lgar.cxx) against the
output files of the reference's own executable: the committed head of its CSVs (tools/make_standalone_golden.py)
and, when oracle/_ref/lasam_standalone travelled to this box, a full run of it."""
import re
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import GOLDEN, ROOT, config_for

EXE = ROOT / "lgar_c_b200" / "lasam_b200_standalone"
REF_EXE = ROOT / "oracle" / "_ref" / "lasam_standalone"
CASE = "Phillipsburg_with_reservoir"


def read_variables(path, nrows=None):
    lines = Path(path).read_text().splitlines()
    header = lines[0].split(",")
    rows = lines[1:] if nrows is None else lines[1:nrows + 1]
    stamps = [r.split(",")[0] for r in rows]
    vals = np.array([[float(x) for x in r.split(",")[1:]] for r in rows])
    return header, stamps, vals


def read_layers(path, nsteps=None):
    lines = Path(path).read_text().splitlines()
    out = []
    for i in range(0, len(lines) - 1, 2):
        assert lines[i].startswith("# Timestep = ")
        fronts = [tuple(float(x) for x in m.split(",")) for m in re.findall(r"\(([^)]*)\)", lines[i + 1])]
        out.append((lines[i], fronts))
        if nsteps is not None and len(out) == nsteps:
            break
    return out


def assert_variables_close(got, want):
    # the files carry 15 decimals of values in metres: 1e-9 relative (BASELINE.json) + the print resolution;
    # column 10 is the mass-balance residual (compared absolutely)
    assert got.shape == want.shape
    err = np.abs(got - want)
    tol = 1e-9 * np.maximum(np.abs(got), np.abs(want)) + 3e-12
    tol[:, 10] = 1e-11
    assert np.all(err <= tol), np.argwhere(err > tol)[:5]


def test_usage_without_arguments():
    """No GPU needed: the executable exists after build() and prints its usage like the reference does (bmi_main_lgar.cxx:39-44)."""
    r = subprocess.run([str(EXE)], capture_output=True, text=True)
    assert r.returncode == 0 and "CONFIGURATION_FILE" in r.stdout


@pytest.mark.gpu
def test_hourly_loop_writes_the_reference_files(cuda_device, tmp_path):
    cfg = config_for(CASE, tmp_path)
    n = 400
    subprocess.run([str(EXE), str(cfg), "--layers", "--out-dir", str(tmp_path)], check=True, capture_output=True)
    hdr, stamps, got = read_variables(tmp_path / "data_variables.csv", n)
    hdr0, stamps0, want = read_variables(GOLDEN / f"standalone_{CASE}_variables_head.csv")
    assert hdr == hdr0 and stamps == stamps0
    assert_variables_close(got, want)
    lg, lw = read_layers(tmp_path / "data_layers.csv", n), read_layers(GOLDEN / f"standalone_{CASE}_layers_head.csv")
    assert len(lg) == len(lw) == n
    for (h1, f1), (h2, f2) in zip(lg, lw):
        assert h1 == h2 and len(f1) == len(f2)           # step header and front count exact
        for a, b in zip(f1, f2):
            assert a[2] == b[2] and a[3] == b[3]        # layer, front number
            assert abs(a[0] - b[0]) <= 2e-6 and abs(a[1] - b[1]) <= 2e-6 and abs(a[4] - b[4]) <= 2e-6 * max(1.0, abs(b[4]))  # "%lf": 6 decimals
    if REF_EXE.exists():  # the whole year against the reference executable run here
        d = tmp_path / "ref"
        d.mkdir()
        subprocess.run([str(REF_EXE), str(cfg)], cwd=d, check=True, stdout=subprocess.DEVNULL)
        _, s_all, want_all = read_variables(d / "data_variables.csv")
        _, g_all_s, got_all = read_variables(tmp_path / "data_variables.csv")
        assert s_all == g_all_s
        assert_variables_close(got_all, want_all)


@pytest.mark.gpu
def test_streamed_batch_run_and_binary_formats(cuda_device, tmp_path):
    """Streamed mode: 5 columns with a circular time shift; column 0 equals the hour-by-hour files, a shifted column
    equals a single-column run on the shifted forcing (through the binary forcing format, f64), the binary output holds
    the same numbers as the CSVs, and an f32 run stays within f32 resolution."""
    cfg = config_for(CASE, tmp_path)
    N, shift = 5, 1000
    a = tmp_path / "a"
    a.mkdir()
    subprocess.run([str(EXE), str(cfg), "--columns", str(N), "--shift", str(shift), "--chunk", "1000", "--write-columns", "0,3", "--out-dir", str(a),
                    "--out-bin", str(a / "out.lgf"), "--out-vars", "total_discharge,soil_storage"], check=True, capture_output=True)
    _, _, want = read_variables(GOLDEN / f"standalone_{CASE}_variables_head.csv")
    _, _, c0 = read_variables(a / "data_variables_c0.csv")
    assert_variables_close(c0[:len(want)], want)
    raw = (a / "out.lgf").read_bytes()
    magic, dtype, nsteps, ncol, nvars, _ = struct.unpack("<4siqqii", raw[:32])
    assert magic == b"LGF1" and dtype == 0 and ncol == N and nvars == 2 and nsteps == len(c0)
    names = [raw[32 + 64 * k: 96 + 64 * k].split(b"\0")[0].decode() for k in range(nvars)]
    assert names == ["total_discharge", "soil_storage"]
    planes = np.frombuffer(raw, dtype="<f8", offset=32 + 64 * nvars).reshape(nvars, nsteps, ncol)
    assert np.allclose(planes[0, :, 0], c0[:, 6], rtol=0, atol=6e-16) and np.allclose(planes[1, :, 0], c0[:, 5], rtol=0, atol=6e-16)
    _, _, c3 = read_variables(a / "data_variables_c3.csv")
    assert np.allclose(planes[1, :, 3], c3[:, 5], rtol=0, atol=6e-16)
    # column 3 alone, fed its shifted forcing through an LGF1 forcing file
    from oracle.oracle_py import DATA, read_forcing
    P, E = read_forcing(DATA / "forcing" / "forcing_data_resampled_uniform_Phillipsburg.csv")
    idx = (np.arange(nsteps) + 3 * shift) % len(P)
    b = tmp_path / "b"
    b.mkdir()
    with open(b / "forcing.lgf", "wb") as f:
        f.write(struct.pack("<4siqqii", b"LGF1", 0, nsteps, 1, 2, 0))
        f.write(b"precipitation_rate".ljust(64, b"\0") + b"potential_evapotranspiration_rate".ljust(64, b"\0"))
        f.write(P[idx].astype("<f8").tobytes() + E[idx].astype("<f8").tobytes())
    subprocess.run([str(EXE), str(cfg), "--forcing-bin", str(b / "forcing.lgf"), "--out-dir", str(b)], check=True, capture_output=True)
    _, _, single = read_variables(b / "data_variables.csv")
    assert np.array_equal(single, c3)
    # f32 forcing and outputs: same run within f32 resolution of the forcing (storage in metres)
    c = tmp_path / "c"
    c.mkdir()
    subprocess.run([str(EXE), str(cfg), "--columns", "2", "--f32", "--out-dir", str(c), "--out-bin", str(c / "o.lgf")], check=True, capture_output=True)
    _, _, f0 = read_variables(c / "data_variables_c0.csv")
    assert np.max(np.abs(f0[:, 5] - c0[:, 5])) < 1e-4
