"""Runs the kernel text (tests/hostsim, built with -fsanitize=address,undefined) through the golden cases in every program
form, and through the 19-front column of BASELINE config 5.  Started by tests/test_hostsim.py::test_kernel_text_under_sanitizers
in a process of its own with the sanitizer runtimes preloaded; any out-of-bounds access, signed overflow, misaligned access or
invalid shift in the physics / staged-program code aborts it.  Test infrastructure."""
import ctypes as C
import pathlib
import sys
import tempfile

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent.parent / "tools"))
from helpers import STANDARD, RANDOM, load_golden, config_for  # noqa: E402
import test_hostsim as th  # noqa: E402

lib = C.CDLL(sys.argv[1])
lib.hostsim_run_series.argtypes = [C.c_char_p, th._ip, th._dp, C.c_int, th._dp, th._dp, th._dp, th._ip, C.c_int, th._dp, th._ip, C.c_int, th._ip, C.c_char_p, C.c_int]
lib.hostsim_run_series.restype = C.c_int
tmp = pathlib.Path(tempfile.mkdtemp())
steps = 0
for name in STANDARD + RANDOM[:4]:
    g = load_golden(name)
    n = min(len(g["precip"]), 2000)
    for force in (0, 1, 2, 3, 4, 5):
        done, scal, nf, fr, fl, st = th.run(lib, config_for(name, tmp), g["precip"][:n], g["pet"][:n], force)
        assert done == n and np.array_equal(nf, g["nfronts"][:n]), (name, force)
        steps += n
# the deep column (19 fronts at hour 5641): capacity-32 arrays and both flag words in use
import torch  # noqa: E402
import parity_sampling_full as pf  # noqa: E402
from oracle.oracle_py import write_config  # noqa: E402
c = pf.Config(5, tmp)
ids = np.array([5588097], dtype=np.int64)
T = 5800
P, E = (x.numpy()[:, 0] for x in c.forcing(torch.from_numpy(ids), 0, T))
cfg = write_config(c.cfg, tmp / "deep.txt", layer_soil_type=",".join(map(str, c.soils(ids)[0])), layer_thickness=",".join(map(str, c.layers(ids)[0])) + "[cm]")
for force in (0, 4):
    done, scal, nf, fr, fl, st = th.run(lib, cfg, P, E, force, cap=32)
    assert done == T and nf.max() == 19
    steps += T
# malformed configurations through the product's host-side reader (lgar_c_b200/csrc/lgar_config.hpp, the file lgar_engine.cu
# parses with): a clean error with a message, never a crash.  (key overrides, expected: None = must run, else text in the message)
from oracle.oracle_py import DATA, absolutize_config  # noqa: E402
base = absolutize_config(DATA / "configs" / "config_lasam_Phillipsburg.txt", out_dir=tmp)
bad = {
    "missing layer_thickness": (dict(layer_thickness=None), "does not set layer_thickness"),
    "missing soil file key": (dict(soil_params_file=None), "does not set soil_params_file"),
    "soil file absent": (dict(soil_params_file="/nonexistent/soils.dat"), "Can't open input file"),
    "more soil types than layers": (dict(layer_soil_type="13,14,15,16"), "differ in length"),
    "fewer soil types than layers": (dict(layer_soil_type="13,14"), "differ in length"),
    "five layers": (dict(layer_thickness="10,10,10,10,10[cm]", layer_soil_type="1,2,3,4,5"), "layers"),
    "soil type beyond the table (invalid soil: precipitation passes through)": (dict(layer_soil_type="13,14,99"), None),
    "soil type zero": (dict(layer_soil_type="0,14,15"), ">= 1"),
    "40 GIUH ordinates": (dict(giuh_ordinates=",".join(["0.025"] * 40)), "GIUH"),
    "no GIUH ordinates": (dict(giuh_ordinates=None), "does not set giuh_ordinates"),
    "malformed number": (dict(initial_psi="abc[cm]"), "initial_psi"),
    "zero timestep": (dict(timestep="0[sec]"), "must be positive"),
    "missing timestep": (dict(timestep=None), "does not set timestep"),
    "missing forcing_resolution": (dict(forcing_resolution=None), "does not set forcing_resolution"),
    "empty file": (None, "does not set"),
}
P = np.array([0.0, 5.0, 20.0, 0.0, 0.0, 1.0] * 4)
E = np.full(24, 0.1)
for k, (name, (rep, expect)) in enumerate(bad.items()):
    path = tmp / f"bad{k}.txt"
    if rep is None:
        path.write_text("")
    else:
        write_config(base, path, **rep)
    n = len(P)
    scal = np.zeros((n, 12)); nf = np.zeros(n, dtype=np.int32); fr = np.zeros((n, 32, 5)); fl = np.zeros((n, 32), dtype=np.int32)
    st = C.c_int(0)
    err = C.create_string_buffer(512)
    done = lib.hostsim_run_series(str(path).encode(), None, None, n, P.ctypes.data_as(th._dp), E.ctypes.data_as(th._dp), scal.ctypes.data_as(th._dp),
                                  nf.ctypes.data_as(th._ip), 32, fr.ctypes.data_as(th._dp), fl.ctypes.data_as(th._ip), 0, C.byref(st), err, 512)
    msg = err.value.decode()
    if expect is None:
        assert done == n, (name, done, msg)
    else:
        assert done < 0 and expect in msg, (name, done, msg)
# malformed soil-parameter files (lgar_read_vG_param_file, lgar.cxx:2675-2760 -> read_soil_file): too few rows is the
# reference's "greater than the number of lines in the soil data file" error; rows that sscanf cannot read inherit the previous
# row (quirk Q2), physically impossible rows run into a per-column status -- nothing may crash
good = (DATA / "vG_default_params.dat").read_text().splitlines()
soil_files = {
    "empty": ("", "number of lines"),
    "header only": (good[0] + "\n", "number of lines"),
    "5 rows only": ("\n".join(good[:6]) + "\n", "number of lines"),
    "row with missing columns": ("\n".join(good[:14] + ['"Broken" 0.05 0.4'] + good[15:]) + "\n", None),
    "non-numeric row": ("\n".join(good[:14] + ['"Broken" a b c d e'] + good[15:]) + "\n", None),
    "n = 1 (m = 0)": ("\n".join(good[:13] + ['"Flat" 0.05 0.40 0.01 1.0 0.5'] + good[14:]) + "\n", None),
    "theta_r > theta_e": ("\n".join(good[:13] + ['"Inv" 0.45 0.40 0.01 1.5 0.5'] + good[14:]) + "\n", None),
    "5 000-character name": ("\n".join(good[:13] + ['"' + "x" * 5000 + '" 0.05 0.40 0.01 1.5 0.5'] + good[14:]) + "\n", None),
}
for k, (name, (text, expect)) in enumerate(soil_files.items()):
    sf = tmp / f"soil{k}.dat"
    sf.write_text(text)
    path = write_config(base, tmp / f"badsoil{k}.txt", soil_params_file=str(sf))
    n = len(P)
    scal = np.zeros((n, 12)); nf = np.zeros(n, dtype=np.int32); fr = np.zeros((n, 32, 5)); fl = np.zeros((n, 32), dtype=np.int32)
    st = C.c_int(0)
    err = C.create_string_buffer(512)
    done = lib.hostsim_run_series(str(path).encode(), None, None, n, P.ctypes.data_as(th._dp), E.ctypes.data_as(th._dp), scal.ctypes.data_as(th._dp),
                                  nf.ctypes.data_as(th._ip), 32, fr.ctypes.data_as(th._dp), fl.ctypes.data_as(th._ip), 0, C.byref(st), err, 512)
    if expect is not None:
        assert done < 0 and expect in err.value.decode(), (name, done, err.value.decode())
print(f"sanitizers clean over {steps} column-steps, {len(bad)} malformed configurations and {len(soil_files)} malformed soil files")
