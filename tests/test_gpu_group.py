"""Several GPUs behind one handle (include/lgarb_group.h, liblgarb_group.so): needs at least two CUDA devices -- the driver's
single-GPU suite skips it, `gpurun --gpus 2 -- python -m pytest tests/test_gpu_group.py -m gpu` runs it (tools/r2_call13.sh)."""
import numpy as np
import pytest

from helpers import load_golden, config_for

pytestmark = pytest.mark.gpu


def test_group_of_devices_equals_one_engine(cuda_device, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two CUDA devices")
    import bench
    import lgar_c_b200
    from lgar_c_b200.group import LgarGroup
    ndev = min(torch.cuda.device_count(), 4)
    ncol, T = 30_001, 120    # ragged last block
    cfg = bench.write_workload_config(tmp_path)
    soils, layers = bench.column_soils(0, ncol), bench.column_layers(0, ncol)
    P, E = bench.host_forcing(ncol, 0, T, 99)
    one = lgar_c_b200.BmiLGARBatch()
    one.initialize(cfg, ncol, 0, soils, layers)
    ref = one.update_steps_host(P, E, ["total_discharge", "soil_storage"])
    want = one.get_global_mass_balance()
    grp = LgarGroup(list(range(ndev)))
    grp.initialize(cfg, ncol, soils, layers)
    sh = grp.shards()
    assert sh[0][0] == 0 and sum(n for _, n in sh) == ncol and all(sh[k][0] + sh[k][1] == sh[k + 1][0] for k in range(ndev - 1))
    out = grp.update_steps_host(P, E, ["total_discharge", "soil_storage"])
    for n in ref:
        assert np.array_equal(out[n], ref[n], equal_nan=True), n
    gm, tot, ms = grp.gather_mass_balance()
    assert np.array_equal(gm, want, equal_nan=True), "gathered balance terms differ from the single engine's"
    assert np.allclose(tot, want.sum(axis=0), rtol=1e-12, atol=1e-9)
    print(f"{ndev} devices, shards {sh}, NCCL gather {ms:.3f} ms")
    grp.finalize()
