"""SFT coupling (reference lgar.cxx:1109-1149 and the frozen_factor call sites): with
sft_coupled=true the soil temperature profile set through SetValue("soil_temperature_profile") scales the hydraulic
conductivity per layer.  Hour by hour against the UNMODIFIED reference (oracle/_ref), one object per column, with
temperatures that sweep the factor through its whole range [0.05, 1] in every layer."""
import numpy as np
import pytest

from helpers import assert_fronts_close, assert_scalars_close, load_golden
from oracle.oracle_py import DATA, absolutize_config, write_config

pytestmark = pytest.mark.gpu

SOIL_Z = "10,20,30,40,50,60,70,80,90,100.0,110.,120,130.,140.,150.,160.,170.,180.,190.,200.0[cm]"   # configs/config_lasam_sft_ngen.txt


def temperatures(ncol, T, ncell):
    """[T][ncol][ncell] K: around freezing, so that exp(-10 (273.15 - T)) moves between 0.05 and 1."""
    t = np.arange(T)[:, None, None]
    c = np.arange(ncol)[None, :, None]
    z = np.arange(ncell)[None, None, :]
    return 273.15 - 0.18 + 0.22 * np.sin(2 * np.pi * (t / 97.0 + c / 5.0)) + 0.012 * z * np.cos(2 * np.pi * t / 211.0 + c)


@pytest.mark.parametrize("closed_form", ["true", "false"])
def test_sft_coupled_columns_against_reference(cuda_device, ref, tmp_path, closed_form):
    import lgar_c_b200
    g = load_golden("Phillipsburg_with_reservoir")
    extra = {"use_closed_form_G": closed_form}
    if closed_form == "false":                     # the shipped SFT configuration: 300 s sub-steps, numeric Geff, ponding
        extra.update(timestep="300[sec]", ponded_depth_max="2[cm]")
    base = write_config(DATA / "configs" / "config_lasam_Phillipsburg_with_reservoir.txt", tmp_path / "sft.txt", sft_coupled="true", soil_z=SOIL_Z, **extra)
    cfg = absolutize_config(base, out_dir=tmp_path)
    ncol, T, ncell = 6, 360, 20
    shifts = np.arange(ncol) * 1100 + 40
    Pm = np.stack([np.roll(g["precip"], -int(s))[:T] for s in shifts], axis=1)
    Em = np.stack([np.roll(g["pet"], -int(s))[:T] for s in shifts], axis=1)
    Tm = temperatures(ncol, T, ncell)
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol)
    assert b.get_grid_size(4) == ncol * ncell
    hs = [ref.create(cfg) for _ in range(ncol)]
    names = lgar_c_b200.OUTPUT_SCALARS
    got = np.zeros((T, 12, ncol))
    want = np.zeros((T, 12, ncol))
    nf_got = np.zeros((T, ncol), dtype=np.int32)
    nf_want = np.zeros((T, ncol), dtype=np.int32)
    last = [None] * ncol
    for t in range(T):
        if t >= 3:      # the first steps run on the zero profile the reference starts with (factor 0.05 everywhere)
            b.set_value("soil_temperature_profile", Tm[t])
        b.set_value("precipitation_rate", Pm[t])
        b.set_value("potential_evapotranspiration_rate", Em[t])
        b.update()
        for k, n in enumerate(names):
            got[t, k] = b.get_value(n)
        nf_got[t] = b.get_value("soil_num_wetting_fronts")
        for c in range(ncol):
            if t >= 3:
                ref.set_array(hs[c], "soil_temperature_profile", Tm[t, c])
            r = ref.run_series(hs[c], Pm[t:t + 1, c], Em[t:t + 1, c], cap=16)
            assert r["done"] == 1, f"reference stopped at step {t}, column {c}"
            want[t, :, c] = r["scalars"][0]
            nf_want[t, c] = r["nfronts"][0]
            last[c] = r
    assert (b.get_status() & lgar_c_b200.STATUS_FAIL_MASK == 0).all()
    assert np.array_equal(nf_got, nf_want)
    fr = b.get_fronts()
    for c in range(ncol):
        assert_scalars_close(got[:, :, c], want[:, :, c], f"column {c}")
        n_end = int(nf_want[-1, c])
        arr = np.stack([fr[k][c] for k in ("depth", "theta", "psi", "K", "dzdt")], axis=1)
        flags = (fr["layer"][c] | (fr["to_bottom"][c] << 8)).astype(np.int32)
        assert_fronts_close(arr, flags, last[c]["fronts"][0], last[c]["flags"][0], n_end, f"column {c} final fronts")
    # the coupling did something: an uncoupled run of the same forcing differs
    base0 = write_config(base, tmp_path / "nosft.txt", sft_coupled="false")
    b0 = lgar_c_b200.BmiLGARBatch()
    b0.initialize(absolutize_config(base0, out_dir=tmp_path), ncol)
    out0 = b0.update_steps_host(np.ascontiguousarray(Pm), np.ascontiguousarray(Em), ["soil_storage"])
    assert np.max(np.abs(out0["soil_storage"] - got[:, 5, :])) > 1e-4


def test_sft_coupled_columns_in_the_window_path(cuda_device, ref, tmp_path):
    """SFT-coupled columns through the MULTI-STEP verbs (the stage kernels; round 1 stepped coupled runs hour by hour only): the
    temperature profile is set once per window of 24 forcing hours -- a temperature series per window -- and the window runs with
    each column's frozen factors carried in its lane record.  Against the unmodified reference fed the same profiles at the same
    hours: per-step scalars, front counts and final fronts bit for bit; and identical to the engine's own hour-by-hour path."""
    import torch
    import lgar_c_b200
    g = load_golden("Phillipsburg_with_reservoir")
    base = write_config(DATA / "configs" / "config_lasam_Phillipsburg_with_reservoir.txt", tmp_path / "sftw.txt", sft_coupled="true", soil_z=SOIL_Z)
    cfg = absolutize_config(base, out_dir=tmp_path)
    ncol, T, ncell, W = 40, 720, 20, 24
    shifts = np.arange(ncol) * 211 + 40
    Pm = np.ascontiguousarray(np.stack([np.roll(g["precip"], -int(s))[:T] for s in shifts], axis=1))
    Em = np.ascontiguousarray(np.stack([np.roll(g["pet"], -int(s))[:T] for s in shifts], axis=1))
    Tm = temperatures(ncol, T, ncell)
    res = {}
    for mode in (3, 0):
        b = lgar_c_b200.BmiLGARBatch()
        b.initialize(cfg, ncol)
        b.set_window_mode(mode)
        dP, dE = torch.from_numpy(Pm).cuda(), torch.from_numpy(Em).cuda()
        sinks = torch.zeros((12, T, ncol), dtype=torch.float64, device="cuda")
        nfs = torch.zeros((T, ncol), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        for t0 in range(0, T, W):
            if t0 > 0:   # the first window runs on the zero profile the reference starts with
                b.set_value("soil_temperature_profile", Tm[t0])
            b.update_steps_device(W, dP[t0].data_ptr(), dE[t0].data_ptr(), ncol,
                                  {n: sinks[k, t0].data_ptr() for k, n in enumerate(lgar_c_b200.OUTPUT_SCALARS)}, ncol, nfs[t0].data_ptr())
        b.synchronize()
        assert (b.get_status() & lgar_c_b200.STATUS_FAIL_MASK == 0).all()
        res[mode] = (sinks.cpu().numpy().transpose(1, 0, 2), nfs.cpu().numpy(), b.get_fronts())
    got, nf, fr = res[3]
    assert np.array_equal(got, res[0][0]) and np.array_equal(nf, res[0][1])
    for c in range(0, ncol, 3):
        h = ref.create(cfg)
        want = np.zeros((T, 12))
        nfw = np.zeros(T, dtype=np.int32)
        for t0 in range(0, T, W):
            if t0 > 0:
                ref.set_array(h, "soil_temperature_profile", Tm[t0, c])
            r = ref.run_series(h, Pm[t0:t0 + W, c], Em[t0:t0 + W, c], cap=16)
            assert r["done"] == W
            want[t0:t0 + W], nfw[t0:t0 + W] = r["scalars"], r["nfronts"]
        assert np.array_equal(nf[:, c], nfw), f"column {c}"
        assert_scalars_close(got[:, :, c], want, f"column {c}")
        n_end = int(nfw[-1])
        arr = np.stack([fr[k][c] for k in ("depth", "theta", "psi", "K", "dzdt")], axis=1)
        flags = (fr["layer"][c] | (fr["to_bottom"][c] << 8)).astype(np.int32)
        assert_fronts_close(arr, flags, r["fronts"][-1], r["flags"][-1], n_end, f"column {c} final fronts")
        ref.destroy(h)
