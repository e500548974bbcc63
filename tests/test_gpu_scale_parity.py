"""Parity sampling at scale: the bench workload (config-5 shard: random soil
triples from the parsed stat_nom table, two layerings, synthetic hourly forcing) on 100 000 columns,
a seeded random subset diffed against the oracle step by step, plus every column the engine froze
with a failure status -- the oracle (= the reference, which would abort() there) must fail at the
same step for the same reason.  Needs a B200."""
import os

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


def _assert_scheduling_variants_agree(tmp_path, variants):
    """Runs the bench workload on 200 000 columns x 2 windows once per (window mode, wave parameters, environment) variant and
    requires every output, status word and front to be bit-identical to the first variant's."""
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
    for mode, wave, env in variants:
        os.environ.update(env)
        try:
            b = lgar_c_b200.BmiLGARBatch()
        finally:
            for k in env:
                del os.environ[k]
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


def test_scheduling_does_not_change_results(cuda_device, tmp_path):
    """A column's trajectory must not depend on how the wavefront scheduler interleaves it with the others: the bench
    workload on 200 000 columns x 2 windows, run with (a) the defaults, (b) no finishing kernel and a tiny search quantum,
    (c) the finishing kernel taking over early -- the shared-memory kernel (lgar_finish_smem.cu, default) with 32, 8 and 2 records
    per warp, a search quantum of 7 and the fast path of the psi searches, and round 1's one-thread-per-slot kernel with and
    without the fast path --, (d) the warp-tile window kernel (mode 1), (e) correction searches on the main stream, (f) the fast
    search path in every round -- every output, status word and front bit-identical."""
    early = (64, 4, 150_000)
    variants = [(3, None, {}), (3, (4, 1, 0), {}), (3, early, {}), (1, None, {}),
                (3, early, {"LGARB_FINISH_SPW": "32", "LGARB_FINISH_FF": "1"}),
                (3, early, {"LGARB_FINISH_SPW": "8", "LGARB_FINISH_QUANTUM": "7"}),
                (3, early, {"LGARB_FINISH_SPW": "2"}),
                (3, early, {"LGARB_FINISH_SMEM": "0"}),
                (3, early, {"LGARB_FINISH_SMEM": "0", "LGARB_FINISH_FF": "1"}),
                (3, None, {"LGARB_WAVE_CORR_ASYNC": "0"}),
                (3, None, {"LGARB_WAVE_FF_BELOW": "1000000000"})]
    _assert_scheduling_variants_agree(tmp_path, variants)


def test_full_size_shard_properties(cuda_device, oracle, tmp_path):
    """BASELINE.json's bench size (one GPU's share of config 5: 1 250 000 columns) for 240 forcing hours, checked through
    properties that do not need the oracle on every column: (a) columns are independent -- 3 000 of them re-run alone
    in a small batch give bit-identical outputs, fronts and balances; (b) the global mass balance of every column
    that was not frozen closes (|error| <= 1e-6 cm, the reference prints ~1e-7 after a year); (c) 400 of the re-run
    columns are diffed against the oracle; (d) frozen columns are rare and are columns the oracle stops on too."""
    import torch
    import bench
    import lgar_c_b200
    ncol, T = 1_250_000, 240
    cfg = bench.write_workload_config(tmp_path)
    soils, layers = bench.column_soils(0, ncol), bench.column_layers(0, ncol)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(20261020)
    u1 = torch.rand((T, ncol), generator=gen, device="cuda", dtype=torch.float64)
    u2 = torch.rand((T, ncol), generator=gen, device="cuda", dtype=torch.float64)
    dP = torch.where(u1 < 0.03, torch.clamp(-2.5 * torch.log1p(-u2), max=170.0), torch.zeros_like(u1))
    del u1, u2
    dE = torch.from_numpy(bench.pet_series(np.arange(T))).cuda()[:, None].expand(T, ncol).contiguous()
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol, 0, soils, layers)
    q = torch.zeros((T, ncol), dtype=torch.float64, device="cuda")
    st_series = torch.zeros((T, ncol), dtype=torch.float64, device="cuda")
    nfs = torch.zeros((T, ncol), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    b.update_steps_device(T, dP.data_ptr(), dE.data_ptr(), ncol, {"total_discharge": q.data_ptr(), "soil_storage": st_series.data_ptr()}, ncol, nfs.data_ptr())
    b.synchronize()
    status = b.get_status()
    frozen = np.nonzero(status & lgar_c_b200.STATUS_FAIL_MASK)[0]
    assert len(frozen) <= ncol // 20000, f"{len(frozen)} frozen columns"
    # (b) every other column closes its global balance
    worst = 0.0
    for c0 in range(0, ncol, 250_000):
        gm = b.get_global_mass_balance(c0, 250_000)
        ok = (status[c0:c0 + 250_000] & lgar_c_b200.STATUS_FAIL_MASK) == 0
        assert np.all(np.isfinite(gm[ok]))
        worst = max(worst, float(np.abs(gm[ok, 15]).max()))
    assert worst <= 1e-6, f"global mass balance error {worst:.3e} cm"
    # (a) independence: a sample (plus the frozen columns) alone in a small batch
    sample = np.unique(np.concatenate([np.random.default_rng(5).choice(ncol, 3000, replace=False), frozen]))
    idx = torch.from_numpy(sample).cuda()
    sP, sE = dP[:, idx].contiguous(), dE[:, idx].contiguous()
    b2 = lgar_c_b200.BmiLGARBatch()
    b2.initialize(cfg, len(sample), 0, soils[sample], layers[sample])
    q2 = torch.zeros((T, len(sample)), dtype=torch.float64, device="cuda")
    s2 = torch.zeros((T, len(sample)), dtype=torch.float64, device="cuda")
    n2 = torch.zeros((T, len(sample)), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    b2.update_steps_device(T, sP.data_ptr(), sE.data_ptr(), len(sample), {"total_discharge": q2.data_ptr(), "soil_storage": s2.data_ptr()}, len(sample), n2.data_ptr())
    b2.synchronize()
    assert torch.equal(n2, nfs[:, idx])
    assert np.array_equal(q2.cpu().numpy(), q[:, idx].cpu().numpy(), equal_nan=True)
    assert np.array_equal(s2.cpu().numpy(), st_series[:, idx].cpu().numpy(), equal_nan=True)
    assert np.array_equal(b2.get_status(), status[sample])
    f1 = b2.get_fronts()
    for c0 in range(0, len(sample), 1000):   # final fronts of the sample inside the big batch
        for j in range(c0, min(c0 + 1000, len(sample)), 97):
            fb = b.get_fronts(int(sample[j]), 1)
            for k in ("depth", "theta", "psi", "K", "dzdt", "layer", "to_bottom", "nfronts"):
                assert np.array_equal(fb[k][0], f1[k][j], equal_nan=True), (k, int(sample[j]))
    # (c), (d) oracle on part of the sample and on every frozen column
    hP, hE = sP.cpu().numpy(), sE.cpu().numpy()
    hq, hn = q2.cpu().numpy(), n2.cpu().numpy()
    pick = np.unique(np.concatenate([np.arange(0, len(sample), max(1, len(sample) // 400)), np.nonzero(np.isin(sample, frozen))[0]]))
    for j in pick:
        c = int(sample[j])
        p = oracle.params_from_config(cfg, layer_soil_type=[int(x) for x in soils[c]], layer_thickness=[float(x) for x in layers[c]])
        s = oracle.new_state(p)
        r = oracle.run_series(p, s, hP[:, j], hE[:, j], cap=16, want_fronts=False)
        done = r["done"]
        if status[c] & lgar_c_b200.STATUS_FAIL_MASK:
            assert done < T and (s.status & 0xFFFF) == (status[c] & 0xFFFF), f"column {c}: status {status[c]:#x} vs oracle {s.status:#x} at {done}"
        else:
            assert done == T, f"column {c}: oracle stopped at {done}"
        assert np.array_equal(hn[:done, j], r["nfronts"][:done]), f"column {c}: front counts differ"
        g, w = hq[:done, j], r["scalars"][:done, 6]
        assert np.all(np.abs(g - w) <= 1e-9 * np.maximum(np.abs(g), np.abs(w)) + 1e-14), f"column {c}: total_discharge differs"


# config-5 columns (global ids) whose front list grows beyond 16 in the first 7400 hours of the year: the oracle of the
# 10 000-column full-duration sample (tools/parity_sampling_full.py) reaches 17-19 fronts there
DEEP_FRONT_COLUMNS = [1598581, 1978823, 5021247, 5588097, 8081421, 8308783, 8471861]


def test_more_than_16_fronts(cuda_device, oracle, tmp_path):
    """A full year of BASELINE config 5 takes 0.2 % of the columns past 16 wetting fronts (19 in a 10 000-column sample):
    the front capacity (LGARB_MAX_FRONTS = 32, two flag words per column) must carry them, step by step as the oracle."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import parity_sampling_full as pf
    c = pf.Config(5, tmp_path)
    ids = np.array(DEEP_FRONT_COLUMNS, dtype=np.int64)
    T = 7400
    hP, hE, scal, nf, status, final, info = pf.run_engine(c, ids, np.arange(len(ids)), T, 1850, 3)
    assert info["frozen_columns_in_batch"] == 0
    assert nf.max() >= 19 and (nf.max(axis=0) > 16).all(), nf.max(axis=0)
    r = pf.diff_against_oracle(c, ids, hP, hE, scal, nf, status, final, 2)
    assert r["front_count_mismatch_columns"] == 0 and r["failure_mismatch_columns"] == 0, r["examples"]
    assert r["final_front_flag_mismatch_columns"] == 0 and r["final_front_field_mismatch_columns"] == 0
    assert max(v["columns_outside_bar"] for v in r["scalars"].values()) == 0, r["examples"]


def _run_config_columns(num, ids, T, tmp_path, counters=True):
    """BASELINE config `num` for the global column ids `ids` over T forcing steps on the GPU: per-step scalars [T][12][n],
    front counts [T][n], the forcing the engine was fed, the engine."""
    import sys
    import torch
    import lgar_c_b200
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent / "tools"))
    import parity_sampling_full as pf
    c = pf.Config(num, tmp_path)
    ids = np.asarray(ids, dtype=np.int64)
    n = len(ids)
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(c.cfg, n, 0, c.soils(ids), c.layers(ids))
    if counters:
        b.set_path_counters(True)
    dP, dE = c.forcing(torch.from_numpy(ids).cuda(), 0, T)
    sinks = torch.zeros((12, T, n), dtype=torch.float64, device="cuda")
    nfs = torch.zeros((T, n), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    b.update_steps_device(T, dP.data_ptr(), dE.data_ptr(), n, {nm: sinks[k].data_ptr() for k, nm in enumerate(lgar_c_b200.OUTPUT_SCALARS)}, n, nfs.data_ptr())
    b.synchronize()
    return c, b, sinks.cpu().numpy().transpose(1, 0, 2), nfs.cpu().numpy(), dP.cpu().numpy(), dE.cpu().numpy()


def _oracle_column(oracle, c, cid, P, E):
    soils, layers = c.soils(np.array([cid])), c.layers(np.array([cid]))
    kw = {}
    if soils is not None:
        kw["layer_soil_type"] = [int(x) for x in soils[0]]
    if layers is not None:
        kw["layer_thickness"] = [float(x) for x in layers[0]]
    p = oracle.params_from_config(c.cfg, **kw)
    s = oracle.new_state(p)
    r = oracle.run_series(p, s, np.ascontiguousarray(P), np.ascontiguousarray(E), cap=32)
    return r, s


# Round 1's full-duration sampling (profiles/r1_parity_sampling_full.json) found, with x^y = exp(y log x) on the device: config 5
# column 8 416 695 splitting its front count from the oracle's at step 2 662, column 9 271 602 outside 1e-9 in AET for 15 steps,
# and 59 config-4 columns outside 1e-9 in surface_runoff (the twelve listed there below).  With the host library's pow restated
# on the device (lgar_pow.cuh) every one of them is bit-identical to the oracle for the whole duration.
R1_CONFIG5_OFFENDERS = [8416695, 9271602, 5588097]   # + the 19-front column
R1_CONFIG4_OFFENDERS = [111221, 196626, 211082, 250681, 252734, 261016, 340747, 462633, 471584, 586152, 587606, 643728]
# columns of the seed-99 sample that take the rare branches (tools/path_scan.py): theta < theta_r deletion, domain-bottom
# crossing with the depth search of a crossing front, saturated-front depth loop
RARE_PATH_COLUMNS = {5: [88278, 96160, 11932], 4: [35301, 419274, 12347], 3: [3085, 29707, 342466, 1192]}


@pytest.mark.parametrize("num,ids,T", [(5, R1_CONFIG5_OFFENDERS + RARE_PATH_COLUMNS[5], 8760), (4, R1_CONFIG4_OFFENDERS + RARE_PATH_COLUMNS[4], 7500),
                                       (3, RARE_PATH_COLUMNS[3], 8333)])
def test_round1_offenders_and_rare_paths_full_duration(cuda_device, oracle, tmp_path, num, ids, T):
    """The columns on which round 1 left north_star's bar, and columns that take the reference's rare branches, over the FULL
    duration of their BASELINE config: every scalar of every step, every front count and the final fronts equal the oracle's
    bit for bit; and the engine's path counters (lgarb_get_path_counters) equal the oracle's, branch by branch -- merge, layer
    crossing, domain-bottom crossing, dry-over-wet, theta < theta_r deletion, saturated-front depth loop, crossing + bottom depth
    search, every exit of the psi search (reference src/lgar.cxx:1519-1537,1687-1723,1875-2307,3063-3086)."""
    c, b, scal, nf, P, E = _run_config_columns(num, ids, T, tmp_path)
    assert (b.get_status() & 0xFFFF == 0).all()
    fr = b.get_fronts()
    want_paths = np.zeros(24, dtype=np.int64)
    for j, cid in enumerate(ids):
        r, s = _oracle_column(oracle, c, cid, P[:, j], E[:, j])
        assert r["done"] == T
        assert np.array_equal(nf[:, j], r["nfronts"]), f"config {num} column {cid}: front count differs at step {int(np.argmax(nf[:, j] != r['nfronts']))}"
        assert_scalars_close(scal[:, :, j], r["scalars"], f"config {num} column {cid}")
        n = int(fr["nfronts"][j])
        arr = np.stack([fr[k][j] for k in ("depth", "theta", "psi", "K", "dzdt")], axis=1)
        flags = (fr["layer"][j] | (fr["to_bottom"][j] << 8)).astype(np.int32)
        assert_fronts_close(arr, flags, r["fronts"][-1], r["flags"][-1], n, f"config {num} column {cid} final fronts")
        want_paths += np.array(s.n_path[:])
    got = b.get_path_counters()
    import lgar_c_b200.bmi as bmi
    for i, name in enumerate(bmi.PATH_NAMES):
        assert got[name] == want_paths[i], f"config {num}: path {name}: engine {got[name]} oracle {want_paths[i]}"
    must = ["merge", "cross_layer", "cross_bottom", "dry_over_wet", "saturated_depth", "cross_layer_bottom", "surficial", "search", "search_converged",
            "corr_search", "cached_step", "active_step"] + (["theta_r_delete"] if num in (3, 5) else [])
    for name in must:
        assert got[name] > 0, f"config {num}: the run never took path {name}"


def test_search_exits_other_than_convergence(cuda_device, oracle, tmp_path):
    """The psi search ends on `|psi - psi_prev| < 1e-15 and factor < 1e-13` (reference src/lgar.cxx:3063) in the synthetic cases
    (3 of 4 900 searches of the golden cases); the engine must take that exit where the oracle does."""
    from helpers import load_golden, config_for
    import lgar_c_b200
    import torch
    want = 0
    got = 0
    for name in ("synth_0", "synth_1", "synth_2", "unittest"):
        g = load_golden(name)
        cfg = config_for(name, tmp_path)
        T = len(g["precip"])
        b = lgar_c_b200.BmiLGARBatch()
        b.initialize(cfg, 1)
        b.set_path_counters(True)
        dP, dE = torch.from_numpy(g["precip"][:, None].copy()).cuda(), torch.from_numpy(g["pet"][:, None].copy()).cuda()
        torch.cuda.synchronize()
        b.update_steps_device(T, dP.data_ptr(), dE.data_ptr(), 1)
        b.synchronize()
        p = oracle.params_from_config(cfg)
        s = oracle.new_state(p)
        oracle.run_series(p, s, g["precip"], g["pet"], want_fronts=False)
        pc = b.get_path_counters()
        assert pc["search_step_exit"] == s.n_path[11] and pc["search"] == s.n_path[9], name
        want += s.n_path[11]
        got += pc["search_step_exit"]
    assert got == want and got > 0
