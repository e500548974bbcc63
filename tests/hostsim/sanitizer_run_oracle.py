"""The plain-C oracle (oracle/lgar_oracle.c, built with -fsanitize=address,undefined and handed over in LGAR_ORACLE_SO)
through every golden case, bit for bit against the unmodified reference's vectors, and through the 19-front column.  Started by
tests/test_hostsim.py::test_kernel_text_under_sanitizers with the sanitizer runtimes preloaded.  Test infrastructure."""
import pathlib
import sys
import tempfile

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent.parent / "tools"))
from helpers import STANDARD, RANDOM, load_golden, config_for  # noqa: E402
from oracle.oracle_py import Oracle, write_config  # noqa: E402

orc = Oracle()
tmp = pathlib.Path(tempfile.mkdtemp())
steps = 0
for name in STANDARD + RANDOM:
    g = load_golden(name)
    p = orc.params_from_config(config_for(name, tmp))
    s = orc.new_state(p)
    r = orc.run_series(p, s, g["precip"], g["pet"], cap=16)
    assert r["done"] == len(g["precip"]) and np.array_equal(r["nfronts"], g["nfronts"]) and np.array_equal(r["scalars"], g["scalars"]), name
    steps += r["done"]
import torch  # noqa: E402
import parity_sampling_full as pf  # noqa: E402
c = pf.Config(5, tmp)
ids = np.array([5588097], dtype=np.int64)
P, E = (x.numpy()[:, 0] for x in c.forcing(torch.from_numpy(ids), 0, 5800))
cfg = write_config(c.cfg, tmp / "deep.txt", layer_soil_type=",".join(map(str, c.soils(ids)[0])), layer_thickness=",".join(map(str, c.layers(ids)[0])) + "[cm]")
p = orc.params_from_config(cfg)
r = orc.run_series(p, orc.new_state(p), P, E, cap=32)
assert r["done"] == 5800 and r["nfronts"].max() == 19
steps += 5800
print(f"oracle: sanitizers clean over {steps} column-steps")
