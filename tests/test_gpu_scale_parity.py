"""Parity sampling at scale: the bench workload (config-5 shard: random soil
triples from the parsed stat_nom table, two layerings, synthetic hourly forcing) on 100 000 columns,
a seeded random subset diffed against the oracle step by step, plus every column the engine froze
with a failure status -- the oracle (= the reference, which would abort() there) must fail at the
same step for the same reason.  Needs a B200."""
import numpy as np
import pytest

from helpers import assert_scalars_close, assert_fronts_close

pytestmark = pytest.mark.gpu


def test_bench_workload_sample_against_oracle(cuda_device, oracle, tmp_path):
    import torch
    import bench
    import lgar_c_b200
    ncol, T, nsample = 100_000, 600, 1200
    cfg = bench.write_workload_config(tmp_path)
    soils, layers = bench.column_soils(0, ncol), bench.column_layers(0, ncol)
    P, E = bench.host_forcing(ncol, 0, T, 4242)
    # make the forcing burstier than the bench's (storms of several hours): more merges, crossings, ponding
    rng = np.random.default_rng(7)
    storm = rng.random((T, ncol)) < 0.004
    for k in range(1, 5):
        P[k:] = np.where(storm[:-k], np.maximum(P[k:], 6.0 * rng.random((T - k, ncol))), P[k:])
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol, 0, soils, layers)
    dP, dE = torch.from_numpy(P).cuda(), torch.from_numpy(E).cuda()
    sample = np.sort(np.random.default_rng(99).choice(ncol, nsample, replace=False))
    sinks = torch.zeros((12, T, ncol), dtype=torch.float64, device="cuda")
    nfs = torch.zeros((T, ncol), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()  # the engine runs on its own stream: inputs/sinks made by torch must be complete
    b.update_steps_device(T, dP.data_ptr(), dE.data_ptr(), ncol, {n: sinks[k].data_ptr() for k, n in enumerate(lgar_c_b200.OUTPUT_SCALARS)},
                          ncol, nfs.data_ptr())
    b.synchronize()
    status = b.get_status()
    failed = np.nonzero(status & lgar_c_b200.STATUS_FAIL_MASK)[0]
    cols = np.unique(np.concatenate([sample, failed[:300]]))
    scal = sinks[:, :, torch.from_numpy(cols).cuda()].cpu().numpy().transpose(1, 0, 2)   # [T][12][n]
    nf = nfs[:, torch.from_numpy(cols).cuda()].cpu().numpy()
    fr = {k: v for k, v in b.get_fronts().items()}
    n_checked_fail = 0
    for j, c in enumerate(cols):
        p = oracle.params_from_config(cfg, layer_soil_type=[int(x) for x in soils[c]], layer_thickness=[float(x) for x in layers[c]])
        s = oracle.new_state(p)
        r = oracle.run_series(p, s, P[:, c], E[:, c], cap=16)
        done = r["done"]
        if status[c] & lgar_c_b200.STATUS_FAIL_MASK:
            # the engine froze this column: the oracle must have stopped at the same step for the same reason
            assert done < T, f"column {c}: engine status {status[c]:#x} but the oracle ran through"
            assert (s.status & 0xFFFF) == (status[c] & 0xFFFF), f"column {c}: status {status[c]:#x} vs oracle {s.status:#x}"
            first_nan = int(np.argmax(np.isnan(scal[:, 5, j])))
            assert first_nan == done, f"column {c}: engine froze at step {first_nan}, oracle at {done}"
            n_checked_fail += 1
        else:
            assert done == T, f"column {c}: oracle stopped at {done} (status {s.status:#x}) but the engine ran through"
        assert np.array_equal(nf[:done, j], r["nfronts"][:done]), f"column {c} soils {soils[c]}: front count differs at step {int(np.argmax(nf[:done, j] != r['nfronts'][:done]))}"
        assert_scalars_close(scal[:done, :, j], r["scalars"][:done], f"column {c} soils {soils[c]}")
        if done == T:
            n = int(fr["nfronts"][c])
            arr = np.stack([fr[k][c] for k in ("depth", "theta", "psi", "K", "dzdt")], axis=1)
            flags = (fr["layer"][c] | (fr["to_bottom"][c] << 8)).astype(np.int32)
            assert_fronts_close(arr, flags, r["fronts"][-1], r["flags"][-1], n, f"column {c} final fronts")
    print(f"checked {len(cols)} columns x {T} steps against the oracle; {len(failed)} of {ncol} columns frozen, {n_checked_fail} of them cross-checked")


def test_scheduling_does_not_change_results(cuda_device, tmp_path):
    """A column's trajectory must not depend on how the wavefront scheduler interleaves it with the
    others: the bench workload on 200 000 columns x 2 windows, run with (a) the defaults, (b) no
    finishing kernel and a tiny search quantum, (c) the finishing kernel taking over early, and
    (d) the warp-tile window kernel (mode 1) -- every output, status word and front bit-identical."""
    import torch
    import bench
    import lgar_c_b200
    ncol, T = 200_000, 96
    cfg = bench.write_workload_config(tmp_path)
    soils, layers = bench.column_soils(0, ncol), bench.column_layers(0, ncol)
    P, E = bench.host_forcing(ncol, 0, T, 777)
    P[10:14] = np.maximum(P[10:14], 8.0)     # a storm on every column: the whole batch active at once
    dP, dE = torch.from_numpy(P).cuda(), torch.from_numpy(E).cuda()
    res = []
    for mode, wave in ((3, None), (4, None), (4, (64, 4, 150_000)), (3, (4, 1, 0)), (3, (64, 4, 150_000)), (1, None)):
        b = lgar_c_b200.BmiLGARBatch()
        b.initialize(cfg, ncol, 0, soils, layers)
        b.set_window_mode(mode)
        if wave is not None:
            b.set_wave_params(*wave)
        q = torch.zeros((T, ncol), dtype=torch.float64, device="cuda")
        nfs = torch.zeros((T, ncol), dtype=torch.int32, device="cuda")
        for t0, n in ((0, 60), (60, T - 60)):
            torch.cuda.synchronize()
            b.update_steps_device(n, dP[t0].data_ptr(), dE[t0].data_ptr(), ncol, {"total_discharge": q[t0].data_ptr()}, ncol, nfs[t0].data_ptr())
        b.synchronize()
        res.append((q.cpu().numpy(), nfs.cpu().numpy(), b.get_status(), b.get_fronts(), b.get_global_mass_balance(),
                    np.stack([b.get_value(n) for n in lgar_c_b200.OUTPUT_SCALARS])))
        del b
    q0, n0, st0, f0, m0, v0 = res[0]
    for (q1, n1, st1, f1, m1, v1) in res[1:]:
        assert np.array_equal(st0, st1), f"status differs on {np.nonzero(st0 != st1)[0][:10]}"
        assert np.array_equal(n0, n1) and np.array_equal(q0, q1, equal_nan=True)
        assert np.array_equal(m0, m1, equal_nan=True) and np.array_equal(v0, v1, equal_nan=True)
        for k in ("depth", "theta", "psi", "K", "dzdt", "layer", "to_bottom", "nfronts"):
            assert np.array_equal(f0[k], f1[k], equal_nan=True), k
