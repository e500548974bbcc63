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
print(f"sanitizers clean over {steps} column-steps")
