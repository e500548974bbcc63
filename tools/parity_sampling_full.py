#!/usr/bin/env python
"""Parity sampling at full duration ("Parity sampling"): for BASELINE configs 3, 4 and 5 a
fixed random subset of columns (default 10 000, seed 99, drawn from the config's whole column range) is run
by the oracle for the FULL duration of the config and every forcing step is diffed against the engine.

    python tools/parity_sampling_full.py [--configs 3,4,5] [--sample 10000] [--batch 250000] [--chunk 240]
                                         [--steps N] [--out gpurun_out/parity_full.json]

The engine runs one batch per config: the sampled columns plus filler columns of the same config (so that every
regime of the window scheduler is exercised: bulk rounds, drop-out rounds, finishing kernel), `chunk` forcing hours
per lgarb_update_steps_device call, all 12 BMI scalars + the front count of every step captured.  A column's
soils / layering / forcing depend on its GLOBAL index in the config only, so the batch composition does not matter
(tests/test_gpu_scale_parity.py::test_full_size_shard_properties checks that independence bit for bit).

  config 3  Bushland flags + layering, 1 M columns, soils i.i.d. from the 12 parsed rows of vG_params_stat_nom_ordered.dat
            (counter-based hash, seed 12345, draws 3c..3c+2), Bushland forcing CSV shifted by (c*7919) mod 8760; 8333 steps
  config 4  Phillipsburg_with_reservoir as shipped (soils 13,14,15), 4 M columns, Phillipsburg CSV with the same shift and
            a per-column precipitation multiplier 0.5 + U(0,1) (seed 2024); 7500 steps
  config 5  the bench workload (bench.py), 10 M columns, synthetic hourly forcing keyed by (column, hour), seed 777; 8760 steps

Test infrastructure: this is one of the places that may execute oracle/.  --selftest replaces the engine by the
unmodified reference (oracle/_ref) on a tiny sample so that the bookkeeping can be exercised without a GPU.
--selfnoise replaces it by the SAME unmodified reference sources compiled with FMA contraction allowed
(oracle/_ref/liblgar_ref_fma.so, `make -C oracle ref_fma`): the deviations it reports are those between two legitimate
builds of the reference itself -- the floor under the 1e-9 bar.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
from oracle.oracle_py import Oracle, SCALAR_NAMES, absolutize_config, read_forcing, write_config  # noqa: E402
from helpers import RTOL  # noqa: E402

# north_star's bar: 1e-9 relative per step, front counts exact.  No absolute floors: a value that should be zero must be zero.
# The one exception is the per-step mass-balance residual, which is rounding noise around zero (~1e-12 cm) and is compared
# absolutely as SURVEY.md 8c prescribes.  Beside the bar the report counts what the engine is built for since round 2: columns
# whose every scalar of every step EQUALS the oracle's bit for bit.
ATOL_M = 0.0
MBAL_ATOL_M = 1e-11

DATA = ROOT / "data"


def soils_for(ids: np.ndarray, seed: int = 12345) -> np.ndarray:
    """bench.column_soils for arbitrary global column ids."""
    out = np.empty((len(ids), 3), dtype=np.int32)
    c = ids.astype(np.uint64)
    with np.errstate(over="ignore"):
        for k in range(3):
            h = bench.splitmix64(c * np.uint64(3) + np.uint64(k) + np.uint64(seed) * np.uint64(0x100000001B3))
            out[:, k] = (h % np.uint64(12)).astype(np.int32) + 1
    return out


def layers_for(ids: np.ndarray) -> np.ndarray:
    out = np.empty((len(ids), 3))
    out[ids % 2 == 0] = bench.PH_LAYERS
    out[ids % 2 == 1] = bench.BU_LAYERS
    return out


class Config:
    """What a BASELINE config means for a set of global column ids."""

    def __init__(self, num: int, tmp: Path):
        self.num = num
        if num == 3:
            self.n_total, self.steps = 1_000_000, 8333
            base = absolutize_config(DATA / "configs" / "config_lasam_Bushland.txt", out_dir=tmp)
            self.cfg = write_config(base, tmp / "config3.txt", soil_params_file=str(DATA / "vG_params_stat_nom_ordered.dat"),
                                    max_valid_soil_types="12", layer_soil_type="1,2,3", forcing_file=None)
            self.csv = read_forcing(DATA / "forcing" / "forcing_data_resampled_uniform_Bushland.csv")
        elif num == 4:
            self.n_total, self.steps = 4_000_000, 7500
            base = absolutize_config(DATA / "configs" / "config_lasam_Phillipsburg_with_reservoir.txt", out_dir=tmp)
            self.cfg = write_config(base, tmp / "config4.txt", forcing_file=None)
            self.csv = read_forcing(DATA / "forcing" / "forcing_data_resampled_uniform_Phillipsburg.csv")
        elif num == 5:
            self.n_total, self.steps = 10_000_000, 8760
            d = tmp / "c5"
            d.mkdir(exist_ok=True)
            self.cfg = bench.write_workload_config(d)
            self.csv = None
        else:
            raise ValueError(num)

    def soils(self, ids):
        return soils_for(ids) if self.num in (3, 5) else None

    def layers(self, ids):
        return layers_for(ids) if self.num == 5 else None

    # forcing of hours [t0, t0+nt) for the columns `ids` (torch int64 tensor on any device): [nt][n] f64 tensors.
    # One definition for the device (engine runs) and the CPU (--selftest); the oracle is always fed the values the
    # engine was fed, copied back from the device.
    def forcing(self, ids_t, t0, nt):
        import torch
        dev = ids_t.device
        t = torch.arange(t0, t0 + nt, dtype=torch.int64, device=dev)
        if self.num in (3, 4):
            if getattr(self, "_csv_dev", None) is None or self._csv_dev[0].device != dev:
                self._csv_dev = (torch.from_numpy(self.csv[0]).to(dev), torch.from_numpy(self.csv[1]).to(dev))
            n = self._csv_dev[0].shape[0]
            shift = (ids_t * 7919) % 8760
            idx = (t[:, None] + shift[None, :]) % n
            P, E = self._csv_dev[0][idx], self._csv_dev[1][idx]
            if self.num == 4:
                mult = 0.5 + _u01(_splitmix64_t(ids_t + _i64(2024 * 0x100000001B3)))
                P = P * mult[None, :]
            return P.contiguous(), E.contiguous()
        return bench.forcing_torch(ids_t, t0, nt)   # the bench workload's forcing: one definition (bench.py), same doubles on host and device


def _i64(x: int) -> int:
    """a 64-bit pattern as the signed value torch's int64 arithmetic wraps to"""
    x &= 0xFFFFFFFFFFFFFFFF
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x, k):
    return (x >> k) & ((1 << (64 - k)) - 1)


def _splitmix64_t(x):
    """bench.splitmix64 on torch int64 tensors (wrapping multiply, logical shifts)"""
    z = x + _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def _u01(h):
    import torch
    return _lsr(h, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


# --------------------------------------------------------------------------------------------------
# oracle workers (fork: the big arrays are inherited, not pickled)
# --------------------------------------------------------------------------------------------------
_G = {}


def _oracle_worker(rng_):
    j0, j1 = rng_
    cfgobj, sample = _G["cfg"], _G["sample"]
    hP, hE, scal, nf, status, fin = _G["hP"], _G["hE"], _G["scal"], _G["nf"], _G["status"], _G["final"]
    T = hP.shape[0]
    orc = Oracle()
    soils, layers = cfgobj.soils(sample[j0:j1]), cfgobj.layers(sample[j0:j1])
    wr, wa = np.zeros(12), np.zeros(12)
    viol = np.zeros(12, dtype=np.int64)          # columns with any step outside the parity bar, per scalar
    viol_steps = np.zeros(12, dtype=np.int64)
    res = dict(nf_mismatch=0, fail_mismatch=0, frozen=0, steps=0, front_field_mismatch=0, flags_mismatch=0, bit_equal=0, final_fronts_bit_equal=0, examples=[])
    for j in range(j0, j1):
        kw = {}
        if soils is not None:
            kw["layer_soil_type"] = [int(x) for x in soils[j - j0]]
        if layers is not None:
            kw["layer_thickness"] = [float(x) for x in layers[j - j0]]
        p = orc.params_from_config(cfgobj.cfg, **kw)
        s = orc.new_state(p)
        r = orc.run_series(p, s, hP[:, j], hE[:, j], want_fronts=False)
        done = r["done"]
        gfail = bool(status[j] & 0xFFFF)
        if gfail:
            res["frozen"] += 1
        if gfail != (done < T) or (gfail and (s.status & 0xFFFF) != (int(status[j]) & 0xFFFF)):
            res["fail_mismatch"] += 1
            res["examples"].append((int(sample[j]), "failure", int(done), int(status[j]), int(s.status)))
        elif gfail:
            first_nan = int(np.argmax(np.isnan(scal[:, 5, j])))
            if first_nan != done:
                res["fail_mismatch"] += 1
                res["examples"].append((int(sample[j]), "failure step", int(done), first_nan))
        res["steps"] += int(done)
        if not np.array_equal(nf[:done, j], r["nfronts"][:done]):
            res["nf_mismatch"] += 1
            res["examples"].append((int(sample[j]), "nfronts", int(np.argmax(nf[:done, j] != r["nfronts"][:done]))))
        g, w = scal[:done, :, j], r["scalars"][:done]
        err = np.abs(g - w)
        mag = np.maximum(np.abs(g), np.abs(w))
        rel = np.where(err > ATOL_M, err / np.maximum(mag, 1e-300), 0.0)
        wr = np.maximum(wr, np.nan_to_num(rel).max(axis=0, initial=0.0))
        wa = np.maximum(wa, np.nan_to_num(err).max(axis=0, initial=0.0))
        tol = RTOL * mag + ATOL_M
        tol[:, 10] = MBAL_ATOL_M
        bad = ~(err <= tol)
        res["bit_equal"] += int(done == T and np.array_equal(g, w, equal_nan=True) and np.array_equal(nf[:done, j], r["nfronts"][:done]))
        viol += bad.any(axis=0)
        viol_steps += bad.sum(axis=0)
        if bad.any() and len(res["examples"]) < 8:
            k = int(np.argmax(bad.any(axis=0)))
            t = int(np.argmax(bad[:, k]))
            res["examples"].append((int(sample[j]), SCALAR_NAMES[k], t, float(g[t, k]), float(w[t, k])))
        if done == T and fin is not None:
            arr, flags = orc.fronts_of(s)
            n = int(fin["nfronts"][j])
            if n != len(flags) or not np.array_equal((fin["layer"][j][:n] | (fin["to_bottom"][j][:n] << 8)).astype(np.int32), flags):
                res["flags_mismatch"] += 1
            else:
                eq = True
                for q, name in enumerate(("depth", "theta", "psi", "K", "dzdt")):
                    a, b_ = fin[name][j][:n], arr[:, q]
                    eq = eq and np.array_equal(a, b_)
                    if not np.all(np.abs(a - b_) <= RTOL * np.maximum(np.abs(a), np.abs(b_))):
                        res["front_field_mismatch"] += 1
                        break
                res["final_fronts_bit_equal"] += int(eq)
    res.update(worst_rel=wr, worst_abs=wa, viol=viol, viol_steps=viol_steps)
    return res


def diff_against_oracle(cfgobj, sample, hP, hE, scal, nf, status, final, cores):
    _G.update(cfg=cfgobj, sample=sample, hP=hP, hE=hE, scal=scal, nf=nf, status=status, final=final)
    ns = len(sample)
    step = max(1, (ns + 4 * cores - 1) // (4 * cores))
    jobs = [(j, min(j + step, ns)) for j in range(0, ns, step)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        parts = pool.map(_oracle_worker, jobs)
    dt = time.perf_counter() - t0
    tot = dict(columns=ns, steps_per_column=int(hP.shape[0]), column_steps_diffed=int(sum(p["steps"] for p in parts)),
               front_count_mismatch_columns=int(sum(p["nf_mismatch"] for p in parts)),
               failure_mismatch_columns=int(sum(p["fail_mismatch"] for p in parts)),
               frozen_columns_in_sample=int(sum(p["frozen"] for p in parts)),
               final_front_flag_mismatch_columns=int(sum(p["flags_mismatch"] for p in parts)),
               final_front_field_mismatch_columns=int(sum(p["front_field_mismatch"] for p in parts)),
               columns_bit_equal_every_scalar_every_step=int(sum(p["bit_equal"] for p in parts)),
               columns_with_bit_equal_final_fronts=int(sum(p["final_fronts_bit_equal"] for p in parts)),
               oracle_seconds=dt, oracle_cores=cores, scalars={}, examples=[e for p in parts for e in p["examples"]][:12])
    wr = np.max([p["worst_rel"] for p in parts], axis=0)
    wa = np.max([p["worst_abs"] for p in parts], axis=0)
    vc = np.sum([p["viol"] for p in parts], axis=0)
    vs = np.sum([p["viol_steps"] for p in parts], axis=0)
    for k, n in enumerate(SCALAR_NAMES):
        tot["scalars"][n] = dict(worst_rel=float(wr[k]), worst_abs_m=float(wa[k]), columns_outside_bar=int(vc[k]), steps_outside_bar=int(vs[k]))
    return tot


# --------------------------------------------------------------------------------------------------
def run_engine(cfgobj: Config, ids: np.ndarray, pos: np.ndarray, T: int, chunk: int, mode: int, device: int = 0):
    """The engine on the batch `ids`; returns forcing, 12 scalars and front counts of the sampled positions `pos`
    for every step, status and final fronts of the sample, and device seconds."""
    import torch
    import lgar_c_b200
    nb, ns = len(ids), len(pos)
    b = lgar_c_b200.BmiLGARBatch()
    dev = torch.device("cuda", device)
    b.initialize(cfgobj.cfg, nb, device, cfgobj.soils(ids), cfgobj.layers(ids))
    b.set_window_mode(mode)
    ext = torch.cuda.ExternalStream(b.get_stream(), device=dev)
    dpos = torch.from_numpy(pos).to(dev)
    sinks = torch.zeros((12, chunk, nb), dtype=torch.float64, device=dev)
    nfs = torch.zeros((chunk, nb), dtype=torch.int32, device=dev)
    hP, hE = np.empty((T, ns)), np.empty((T, ns))
    scal = np.empty((T, 12, ns))
    nf = np.empty((T, ns), dtype=np.int32)
    dev_ms = 0.0
    dids = torch.from_numpy(ids).to(dev)
    for t0 in range(0, T, chunk):
        nt = min(chunk, T - t0)
        dP, dE = cfgobj.forcing(dids, t0, nt)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        b.update_steps_device(nt, dP.data_ptr(), dE.data_ptr(), nb, {n: sinks[k].data_ptr() for k, n in enumerate(lgar_c_b200.OUTPUT_SCALARS)},
                              nb, nfs.data_ptr())
        e1.record(ext)
        b.synchronize()
        e1.synchronize()
        dev_ms += e0.elapsed_time(e1)
        hP[t0:t0 + nt] = dP[:, dpos].cpu().numpy()
        hE[t0:t0 + nt] = dE[:, dpos].cpu().numpy()
        scal[t0:t0 + nt] = sinks[:, :nt][:, :, dpos].cpu().numpy().transpose(1, 0, 2)
        nf[t0:t0 + nt] = nfs[:nt][:, dpos].cpu().numpy()
        del dP, dE
    status_all = b.get_status()
    fr = b.get_fronts()
    final = {k: v[pos] for k, v in fr.items()}
    gm = b.get_global_mass_balance()
    ok = (status_all & 0xFFFF) == 0
    info = dict(batch_columns=nb, frozen_columns_in_batch=int(np.count_nonzero(~ok)),
                worst_global_mass_balance_cm=float(np.abs(gm[ok, 15]).max()), device_seconds=dev_ms * 1e-3,
                column_steps_per_s=nb * T / (dev_ms * 1e-3))
    return hP, hE, scal, nf, status_all[pos], final, info


def _ref_worker(rng_):
    """columns [j0, j1) of the sample through one build of the unmodified reference"""
    from oracle.oracle_py import RefHarness
    j0, j1 = rng_
    cfgobj, sample, hP, hE, so = _G["cfg"], _G["sample"], _G["hP"], _G["hE"], _G["ref_so"]
    ref = RefHarness(so)
    T = hP.shape[0]
    n_ = j1 - j0
    scal = np.empty((T, 12, n_))
    nf = np.empty((T, n_), dtype=np.int32)
    final = {k: np.zeros((n_, 32)) for k in ("depth", "theta", "psi", "K", "dzdt")}
    final.update(layer=np.zeros((n_, 32), dtype=np.int32), to_bottom=np.zeros((n_, 32), dtype=np.int32), nfronts=np.zeros(n_, dtype=np.int32))
    soils, layers = cfgobj.soils(sample[j0:j1]), cfgobj.layers(sample[j0:j1])
    tmp = Path(tempfile.mkdtemp())
    for j in range(n_):
        rep = {}
        if soils is not None:
            rep["layer_soil_type"] = ",".join(map(str, soils[j]))
        if layers is not None:
            rep["layer_thickness"] = ",".join(map(str, layers[j])) + "[cm]"
        h = ref.create(write_config(cfgobj.cfg, tmp / f"c{j0 + j}.txt", **rep))
        r = ref.run_series(h, hP[:, j0 + j], hE[:, j0 + j], cap=32)
        assert r["done"] == T
        scal[:, :, j], nf[:, j] = r["scalars"], r["nfronts"]
        n = int(r["nfronts"][-1])
        final["nfronts"][j] = n
        for q, name in enumerate(("depth", "theta", "psi", "K", "dzdt")):
            final[name][j, :n] = r["fronts"][-1][:n, q]
        final["layer"][j, :n] = r["flags"][-1][:n] & 0xFF
        final["to_bottom"][j, :n] = r["flags"][-1][:n] >> 8
        ref.destroy(h)
    return scal, nf, final


def run_selftest(cfgobj: Config, sample: np.ndarray, T: int, so=None, cores=1):
    """No GPU: the 'engine' is the unmodified reference (oracle/_ref), or with `so` another build of it."""
    import torch
    ns = len(sample)
    hP, hE = (x.numpy() for x in cfgobj.forcing(torch.from_numpy(sample), 0, T))
    _G.update(cfg=cfgobj, sample=sample, hP=hP, hE=hE, ref_so=so)
    step = max(1, (ns + 4 * cores - 1) // (4 * cores))
    jobs = [(j, min(j + step, ns)) for j in range(0, ns, step)]
    with mp.get_context("fork").Pool(cores) as pool:
        parts = pool.map(_ref_worker, jobs)
    scal = np.concatenate([p[0] for p in parts], axis=2)
    nf = np.concatenate([p[1] for p in parts], axis=1)
    final = {k: np.concatenate([p[2][k] for p in parts], axis=0) for k in parts[0][2]}
    name = "oracle/_ref (selftest)" if so is None else f"unmodified reference, other build: {Path(so).name}"
    return hP, hE, scal, nf, np.zeros(ns, dtype=np.int32), final, dict(batch_columns=ns, engine=name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--sample", type=int, default=10_000)
    ap.add_argument("--batch", type=int, default=250_000)
    ap.add_argument("--chunk", type=int, default=240)
    ap.add_argument("--steps", type=int, default=0, help="0: the config's full duration")
    ap.add_argument("--mode", type=int, default=3)
    ap.add_argument("--cores", type=int, default=0)
    ap.add_argument("--selftest", action="store_true")
    ap.add_argument("--selfnoise", action="store_true",
                    help="no GPU: diff the reference built with FMA contraction (make -C oracle ref_fma) against the oracle -- how far two "
                         "legitimate builds of the reference itself drift apart on the same sample")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "parity_full.json"))
    a = ap.parse_args()
    cores = a.cores or bench.host_cores()
    tmp = Path(tempfile.mkdtemp(prefix="lgar_parity_"))
    report = dict(tool="tools/parity_sampling_full.py", sample_seed=99, bar=dict(rtol=RTOL, atol_m=ATOL_M,
                  mass_balance_atol_m=MBAL_ATOL_M, front_counts="exact"), configs={})
    for num in [int(x) for x in a.configs.split(",")]:
        c = Config(num, tmp)
        T = a.steps or c.steps
        sample = np.sort(np.random.default_rng(99).choice(c.n_total, a.sample, replace=False)).astype(np.int64)
        if a.selftest or a.selfnoise:
            so = ROOT / "oracle" / "_ref" / "liblgar_ref_fma.so" if a.selfnoise else None
            hP, hE, scal, nf, status, final, info = run_selftest(c, sample, T, so, cores)
        else:
            filler = np.random.default_rng(100 + num).choice(c.n_total, max(0, a.batch - a.sample), replace=False).astype(np.int64)
            ids = np.unique(np.concatenate([sample, filler]))
            pos = np.searchsorted(ids, sample)
            hP, hE, scal, nf, status, final, info = run_engine(c, ids, pos, T, a.chunk, a.mode)
        r = diff_against_oracle(c, sample, hP, hE, scal, nf, status, final, cores)
        r["engine"] = info
        r["config_columns"], r["steps"] = c.n_total, T
        report["configs"][f"config{num}"] = r
        worst = max(v["worst_rel"] for v in r["scalars"].values())
        print(f"config {num}: {r['columns']} columns x {T} steps = {r['column_steps_diffed']} column-steps diffed; front-count mismatches "
              f"{r['front_count_mismatch_columns']}, failure mismatches {r['failure_mismatch_columns']}, columns outside the bar "
              f"{max(v['columns_outside_bar'] for v in r['scalars'].values())}, worst rel {worst:.2e}; oracle {r['oracle_seconds']:.1f} s on {cores} cores; engine {info}",
              flush=True)
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text(json.dumps(report, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
