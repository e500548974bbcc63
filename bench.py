#!/usr/bin/env python
"""bench.py -- column-timesteps/second of the batched LGAR/CASAM Update() path.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # the B200 engine
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU implementation

Workload (config.workload = "config5-shard"): BASELINE.json config 5 cut to one GPU's share --
1 250 000 columns per GPU (10 M over 8), 3 layers, layering Phillipsburg / Bushland by global
column parity, soils drawn per column and layer from the PARSED data/vG_params_stat_nom_ordered.dat
(12 rows, reader quirk Q2 included), flags of config_lasam_Phillipsburg_with_reservoir.txt (CASAM:
LGAR + nonlinear reservoir + GIUH, closed-form G, adaptive sub-step, flux caching), synthetic hourly
forcing: wet hour with probability 0.03, intensity Exp(mean 2.5 mm/h) capped at 170, PET =
0.21*max(0,1+cos(2pi(h-14)/24))*(1+0.5 sin(2pi d/365)) mm/h.

A bench "step" is two simulated days: --hours-per-step (48) consecutive Update() calls over every
column of the rank; the timed region is ONE multi-step call of steps x 48 hours (throughput grows with
the length of the call, DESIGN.md section 7: 2.8e8 at 48 h, 4.2e8 at 240 h, 4.9e8 at 480 h).  Before warm-up every column is spun up for --spinup-hours so that the timed
region sees a developed wetting-front population rather than the 3-front initial state.

value   = column-timesteps / s, device-timed (CUDA events on the engine's stream), forcing resident
          in HBM, whole job (sum over ranks / max time over ranks)
e2e     = the same through the BMI verbs with HOST buffers: per forcing hour SetValue(precip),
          SetValue(PET) from pinned host memory, Update(), GetValue(total_discharge) to host
roofline, cpu_baseline: see DESIGN.md section 7.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
DATA = ROOT / "data"

PH_LAYERS = (44.0, 131.0, 25.0)   # data/configs/config_lasam_Phillipsburg.txt
BU_LAYERS = (18.0, 76.0, 135.0)   # data/configs/config_lasam_Bushland.txt
N_GIUH = 5
# Measured DRAM traffic and stage shares come from a committed ncu capture (profiles/r2_window_profile.json), keyed to the sha256
# of the sources of the liblgarb.so it was taken with: a line printed by other sources says so instead of repeating stale numbers.
WINDOW_PROFILE = ROOT / "profiles" / "r2_window_profile.json"


# --------------------------------------------------------------------------------------------------
# workload definition (shared by both arms)
# --------------------------------------------------------------------------------------------------
def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    return z ^ (z >> np.uint64(31))


def column_soils(col0: int, ncol: int, seed: int = 12345) -> np.ndarray:
    """soil index in 1..12 for (global column c, layer k): counter-based so any shard can draw its own."""
    c = np.arange(col0, col0 + ncol, dtype=np.uint64)
    out = np.empty((ncol, 3), dtype=np.int32)
    with np.errstate(over="ignore"):
        for k in range(3):
            h = splitmix64(c * np.uint64(3) + np.uint64(k) + np.uint64(seed) * np.uint64(0x100000001B3))
            out[:, k] = (h % np.uint64(12)).astype(np.int32) + 1
    return out


def column_layers(col0: int, ncol: int) -> np.ndarray:
    c = np.arange(col0, col0 + ncol)
    out = np.empty((ncol, 3))
    out[c % 2 == 0] = PH_LAYERS
    out[c % 2 == 1] = BU_LAYERS
    return out


def write_workload_config(tmpdir: Path) -> Path:
    """config_lasam_Phillipsburg_with_reservoir.txt with the stat_nom soil table, absolute paths."""
    lines = []
    for line in (DATA / "configs" / "config_lasam_Phillipsburg_with_reservoir.txt").read_text().splitlines():
        if line.startswith("soil_params_file="):
            line = f"soil_params_file={DATA / 'vG_params_stat_nom_ordered.dat'}"
        elif line.startswith("max_valid_soil_types="):
            line = "max_valid_soil_types=12"
        elif line.startswith("layer_soil_type="):
            line = "layer_soil_type=1,2,3"
        elif line.startswith("forcing_file="):
            continue
        lines.append(line)
    p = tmpdir / "bench_config5.txt"
    p.write_text("\n".join(lines) + "\n")
    return p


def pet_series(hours: np.ndarray) -> np.ndarray:
    h = hours % 24
    d = (hours // 24) % 365
    return 0.21 * np.maximum(0.0, 1.0 + np.cos(2 * np.pi * (h - 14) / 24.0)) * (1.0 + 0.5 * np.sin(2 * np.pi * d / 365.0))


# ---- the synthetic forcing of config 5, ONE definition for both arms and for the parity tools ---------------------------------
# Counter based: hour h of global column c draws two 64-bit hashes keyed by (c, h); the first decides wet / dry (probability
# 0.03), the second picks the intensity from a 65 536-entry table of the Exp(mean 2.5 mm/h) quantile function capped at 170 mm/h.
# The table is made once with numpy on the host in either arm and only INDEXED on the device, so the device (torch int64
# arithmetic) and the host (numpy uint64) produce the same doubles bit for bit -- no transcendental is evaluated twice.
_FORCING_C = 777 * 0x100000001B3


def exp_intensity_table() -> np.ndarray:
    k = np.arange(65536, dtype=np.float64)
    return np.minimum(-2.5 * np.log1p(-(k + 0.5) / 65536.0), 170.0)


def forcing_np(ids: np.ndarray, hour0: int, nhours: int):
    """[nhours][len(ids)] precipitation and PET (mm/h) of the global columns `ids` for hours [hour0, hour0 + nhours)"""
    t = np.arange(hour0, hour0 + nhours, dtype=np.uint64)
    c = np.asarray(ids).astype(np.uint64)
    with np.errstate(over="ignore"):
        key = (c[None, :] * np.uint64(1000003) + t[:, None]) * np.uint64(2) + np.uint64(_FORCING_C & 0xFFFFFFFFFFFFFFFF)
        h1 = splitmix64(key)
        h2 = splitmix64(key + np.uint64(1))
    wet = (h1 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) < 0.03
    P = np.where(wet, exp_intensity_table()[(h2 >> np.uint64(48)).astype(np.int64)], 0.0)
    E = np.repeat(pet_series(np.arange(hour0, hour0 + nhours))[:, None], len(c), axis=1)
    return np.ascontiguousarray(P), np.ascontiguousarray(E)


def _i64(x: int) -> int:
    x &= 0xFFFFFFFFFFFFFFFF
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x, k):
    return (x >> k) & ((1 << (64 - k)) - 1)


def _splitmix64_t(x):
    """splitmix64 on torch int64 tensors (wrapping multiply, logical shifts)"""
    z = x + _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


_LUT_DEV = {}


def forcing_torch(ids_t, hour0: int, nhours: int):
    """forcing_np on the device of the int64 tensor `ids_t`: the same doubles"""
    import torch
    dev = ids_t.device
    if dev not in _LUT_DEV:
        _LUT_DEV[dev] = torch.from_numpy(exp_intensity_table()).to(dev)
    t = torch.arange(hour0, hour0 + nhours, dtype=torch.int64, device=dev)
    key = (ids_t[None, :] * 1000003 + t[:, None]) * 2 + _i64(_FORCING_C)
    h1 = _splitmix64_t(key)
    h2 = _splitmix64_t(key + 1)
    wet = _lsr(h1, 11).to(torch.float64) * (1.0 / 9007199254740992.0) < 0.03
    P = torch.where(wet, _LUT_DEV[dev][_lsr(h2, 48)], torch.zeros((), dtype=torch.float64, device=dev))
    pet = torch.from_numpy(pet_series(np.arange(hour0, hour0 + nhours))).to(dev)
    return P.contiguous(), pet[:, None].expand(nhours, ids_t.shape[0]).contiguous()


def host_forcing(ncol: int, hour0: int, nhours: int, seed: int):
    """numpy version of round 1's forcing (a numpy generator per worker): kept for tests/test_gpu_scale_parity.py"""
    rng = np.random.default_rng([777, seed, hour0])
    u1 = rng.random((nhours, ncol))
    u2 = rng.random((nhours, ncol))
    P = np.where(u1 < 0.03, np.minimum(-2.5 * np.log1p(-u2), 170.0), 0.0)
    E = np.repeat(pet_series(np.arange(hour0, hour0 + nhours))[:, None], ncol, axis=1)
    return np.ascontiguousarray(P), np.ascontiguousarray(E)


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores
# --------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process = one core: owns `ncol` reference columns, advances them `spin + warm + timed` hours,
    returns wall seconds of the timed part."""
    kind, wid, ncol, col0, spin_h, blocks, cfg_dir = args
    import ctypes as C
    from oracle.oracle_py import Oracle, RefHarness
    soils = column_soils(col0, ncol)
    layers = column_layers(col0, ncol)
    base = Path(cfg_dir) / "bench_config5.txt"  # written once by the parent
    times = []
    if kind == "reference":
        ref = RefHarness()
        hs = []
        made = {}
        for c in range(ncol):
            key = (tuple(soils[c]), tuple(layers[c]))
            if key not in made:
                txt = base.read_text().replace("layer_soil_type=1,2,3", "layer_soil_type=" + ",".join(map(str, soils[c])))
                txt = "\n".join(("layer_thickness=" + ",".join(map(str, layers[c])) + "[cm]") if l.startswith("layer_thickness=") else l
                                for l in txt.splitlines()) + "\n"
                p = Path(cfg_dir) / f"w{wid}_{len(made)}.txt"
                p.write_text(txt)
                made[key] = p
            hs.append(ref.create(made[key]))
        arr = (C.c_void_p * ncol)(*hs)
        sumq = C.c_double(0.0)

        ref.lib.lgref_time_columns_series.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        ref.lib.lgref_time_columns_series.restype = C.c_double

        def advance(P, E):
            return ref.lib.lgref_time_columns_series(arr, ncol, P.shape[0], P.ctypes.data_as(C.POINTER(C.c_double)),
                                                     E.ctypes.data_as(C.POINTER(C.c_double)), C.byref(sumq))
    else:
        orc = Oracle()
        ps, ss = [], []
        for c in range(ncol):
            p = orc.params_from_config(base, layer_soil_type=[int(x) for x in soils[c]], layer_thickness=[float(x) for x in layers[c]])
            ps.append(p)
            ss.append(orc.new_state(p))

        def advance(P, E):
            t0 = time.perf_counter()
            for c in range(ncol):
                orc.lib.lgo_run_series(C.byref(ps[c]), C.byref(ss[c]), P.shape[0], np.ascontiguousarray(P[:, c]).ctypes.data_as(C.POINTER(C.c_double)),
                                       np.ascontiguousarray(E[:, c]).ctypes.data_as(C.POINTER(C.c_double)), None, None, 0, None, None)
            return time.perf_counter() - t0
    ids = np.arange(col0, col0 + ncol, dtype=np.int64)   # the same global columns, hours and forcing as the B200 arm
    hour = 0
    if spin_h:
        P, E = forcing_np(ids, hour, spin_h)
        advance(P, E)
        hour += spin_h
    for nh, timed in blocks:
        P, E = forcing_np(ids, hour, nh)
        dt = advance(P, E)
        hour += nh
        if timed:
            times.append(dt)
    return times


def run_cpu(kind: str, cores: int, cols_per_core: int, spin_h: int, blocks):
    """blocks: list of (hours, timed?).  Returns per-timed-block wall seconds = max over workers."""
    import multiprocessing as mp
    tmp = tempfile.mkdtemp(prefix="lgar_cpu_")
    write_workload_config(Path(tmp))
    ctx = mp.get_context("fork")
    jobs = [(kind, w, cols_per_core, w * cols_per_core, spin_h, blocks, tmp) for w in range(cores)]
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    per_block = np.max(np.array(res), axis=0)
    return per_block


def _exe_worker(args):
    """One process = one core: runs the reference's standalone executable once per column of its share (config file +
    forcing CSV in, data_variables.csv / data_layers.csv out, as the executable does).  Returns wall seconds of the runs."""
    wid, ncol, col0, hours, cfg_dir = args
    exe = ROOT / "oracle" / "_ref" / "lasam_standalone"
    soils = column_soils(col0, ncol)
    layers = column_layers(col0, ncol)
    base = (Path(cfg_dir) / "bench_config5.txt").read_text()
    P, E = forcing_np(np.arange(col0, col0 + ncol, dtype=np.int64), 0, hours)
    stamp = (np.datetime64("2016-10-01T00:00:00") + np.arange(hours).astype("timedelta64[h]")).astype(str)
    total = 0.0
    for c in range(ncol):
        d = Path(cfg_dir) / f"exe_w{wid}_{c}"
        d.mkdir()
        with open(d / "forcing.csv", "w") as f:
            f.write("Time,P(mm/h),PET(mm/h)\n")
            for t in range(hours):
                f.write(f"{stamp[t].replace('T', ' ')},{float(P[t, c])!r},{float(E[t, c])!r}\n")
        txt = base.replace("layer_soil_type=1,2,3", "layer_soil_type=" + ",".join(map(str, soils[c])))
        lines = [("layer_thickness=" + ",".join(map(str, layers[c])) + "[cm]") if l.startswith("layer_thickness=")
                 else (f"endtime={hours}[hr]" if l.startswith("endtime=") else l) for l in txt.splitlines()]
        lines.append(f"forcing_file={d / 'forcing.csv'}")
        (d / "config.txt").write_text("\n".join(lines) + "\n")
        t0 = time.perf_counter()
        r = subprocess.run([str(exe), str(d / "config.txt")], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        total += time.perf_counter() - t0
        if r.returncode != 0 or not (d / "data_variables.csv").exists():
            return -1.0
    return total


def run_standalone_exe(cores: int, cols_per_core: int, hours: int):
    """BASELINE.md section 4, baseline A (what north_star words): the reference's unmodified standalone executable run
    column-parallel as independent processes, one per host core.  Returns column-timesteps/s over all cores, or None."""
    import multiprocessing as mp
    import shutil
    if not (ROOT / "oracle" / "_ref" / "lasam_standalone").exists():
        return None
    tmp = tempfile.mkdtemp(prefix="lgar_exe_")
    try:
        write_workload_config(Path(tmp))
        jobs = [(w, cols_per_core, w * cols_per_core, hours, tmp) for w in range(cores)]
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_exe_worker, jobs)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if min(res) <= 0:
        return None
    return cores * cols_per_core * hours / max(res)


def standalone_exe_entry(cores: int, cols_per_core: int = 32, hours: int = 2880):
    """cpu_baseline.standalone_executable: baseline A of BASELINE.md section 4 beside the in-process harness (baseline B)."""
    try:
        v = run_standalone_exe(cores, cols_per_core, hours)
    except Exception as ex:  # a reported extra, never fatal
        print(f"bench.py: standalone executable baseline failed: {ex}", file=sys.stderr)
        v = None
    if v is None:
        return None
    return {"value": v, "unit": "column-timesteps/s", "cores": cores, "kind": "reference",
            "sample": f"oracle/_ref/lasam_standalone (unmodified src/bmi_main_lgar.cxx), {cores} processes x {cols_per_core} runs x {hours} h, "
                      "one column per run with its own config file and forcing CSV, data_variables.csv + data_layers.csv written "
                      "(file I/O included: that is what the executable does)"}


def host_memory_bytes() -> int:
    """memory this process tree may use: the cgroup limit if there is one, else what the host reports as available"""
    lim = None
    for f in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = Path(f).read_text().strip()
            if v.isdigit() and int(v) < (1 << 60):
                lim = int(v)
                break
        except OSError:
            pass
    try:
        import psutil
        avail = int(psutil.virtual_memory().available)
    except Exception:
        avail = 64 << 30
    return min(lim, avail) if lim else avail


def cpu_kind() -> str:
    return "reference" if (ROOT / "oracle" / "_ref" / "liblgar_ref.so").exists() else "port"


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for i, nm in enumerate(names):
                if len(s) > 2 + i and s[2 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# pieces of the B200 arm
# --------------------------------------------------------------------------------------------------
def lib_sha256() -> str:
    import hashlib
    from lgar_c_b200.bmi import library_path
    return hashlib.sha256(Path(library_path()).read_bytes()).hexdigest()


def kernel_source_sha256() -> str:
    """sha256 over the sources liblgarb.so is built from (two builds of the same sources differ in their bytes: nvcc's fatbin ids),
    the key of the committed ncu capture"""
    import hashlib
    h = hashlib.sha256()
    csrc = ROOT / "lgar_c_b200" / "csrc"
    for f in sorted(list(csrc.glob("*.cu")) + list(csrc.glob("*.cuh")) + list(csrc.glob("*.h")) + list(csrc.glob("*.hpp")) + [csrc / "Makefile"]):
        if f.name in ("lgar_group.cu", "lasam_b200_standalone.cpp"):
            continue
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def committed_window_profile():
    """(traffic bytes per column-timestep, stage time shares, source) of the committed ncu capture if it belongs to this build"""
    try:
        p = json.loads(WINDOW_PROFILE.read_text())
    except Exception:
        return None, None, "no committed capture (profiles/r2_window_profile.json)"
    if p.get("source_sha256") != kernel_source_sha256():
        return None, None, f"the committed capture ({WINDOW_PROFILE.name}) was taken with other sources of liblgarb.so"
    return p.get("dram_bytes_per_column_timestep"), p.get("stage_time_share"), f"{WINDOW_PROFILE.name}: {p.get('source', '')}"


def timed_call(b, ext, nh, dP, dE, stride):
    """one lgarb_update_steps_device call of nh hours, device-timed on the engine's stream; returns milliseconds"""
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    b.update_steps_device(nh, dP.data_ptr(), dE.data_ptr(), stride)
    e1.record(ext)
    e1.synchronize()
    return e0.elapsed_time(e1)


def parity_block(cfgobj, ids, T, chunk, dev_index, cores, mode, what):
    """The engine on the global columns `ids` of a BASELINE config for T forcing steps from the initial state, all 12 BMI scalars
    and the front count of every step captured, diffed against the oracle on the host cores (tools/parity_sampling_full.py:
    one of the places that may execute oracle/ -- as the checker).  Returns the JSON `parity` block."""
    sys.path.insert(0, str(ROOT / "tools"))
    import parity_sampling_full as pf
    pos = np.arange(len(ids))
    hP, hE, scal, nf, status, final, info = pf.run_engine(cfgobj, ids, pos, T, chunk, mode, dev_index)
    r = pf.diff_against_oracle(cfgobj, ids, hP, hE, scal, nf, status, final, cores)
    worst = max(v["worst_rel"] for v in r["scalars"].values())
    return {"what": what, "columns": int(len(ids)), "steps_per_column": int(T), "column_steps_diffed": r["column_steps_diffed"],
            "front_count_mismatch_columns": r["front_count_mismatch_columns"], "failure_mismatch_columns": r["failure_mismatch_columns"],
            "columns_outside_1e-9": int(max(v["columns_outside_bar"] for v in r["scalars"].values())),
            "columns_bit_equal_every_scalar_every_step": r["columns_bit_equal_every_scalar_every_step"],
            "columns_with_bit_equal_final_fronts": r["columns_with_bit_equal_final_fronts"], "worst_rel": worst,
            "frozen_columns_in_sample": r["frozen_columns_in_sample"], "oracle_cores": cores, "oracle_seconds": r["oracle_seconds"],
            "bar": "1e-9 relative per step, no absolute floor (mass_balance residual: 1e-11 m absolute), front counts exact"}


def run_config_workload(args, emit, rank, world, local_rank):
    """--workload config3|config4|config5: the BASELINE config at its stated size over its full duration, this rank owning a
    contiguous block of its columns: one timed pass (no output capture, forcing of a chunk resident before its call is timed) and
    one parity pass over the seed-99 sample of the config's column range that falls into this rank's block."""
    import torch
    import lgar_c_b200
    sys.path.insert(0, str(ROOT / "tools"))
    import parity_sampling_full as pf
    num = int(args.workload[-1])
    tmp = Path(tempfile.mkdtemp(prefix="lgar_bench_"))
    c = pf.Config(num, tmp)
    n_total = args.columns_total or c.n_total
    T = args.hours or c.steps
    per = (n_total + world - 1) // world
    col0, col1 = rank * per, min(n_total, (rank + 1) * per)
    ids = np.arange(col0, col1, dtype=np.int64)
    ncol = len(ids)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(c.cfg, ncol, local_rank, c.soils(ids), c.layers(ids))
    b.set_window_mode(args.mode)
    ext = torch.cuda.ExternalStream(b.get_stream(), device=dev)
    dids = torch.from_numpy(ids).to(dev)
    chunk = args.chunk_hours
    sampler = ClockSampler(local_rank)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = b.get_launch_count()
    ms = 0.0
    for t0 in range(0, T, chunk):
        nt = min(chunk, T - t0)
        dP, dE = c.forcing(dids, t0, nt)      # generated on the device, resident before the timed call
        torch.cuda.synchronize()
        ms += timed_call(b, ext, nt, dP, dE, ncol)
        del dP, dE
    sampler.stop_flag.set()
    launches = b.get_launch_count() - launches0
    st = b.get_status()
    n_failed = int(np.count_nonzero(st & lgar_c_b200.STATUS_FAIL_MASK))
    gm = b.get_global_mass_balance()
    ok = (st & lgar_c_b200.STATUS_FAIL_MASK) == 0
    worst_gmb = float(np.abs(gm[ok, 15]).max()) if ok.any() else float("nan")
    b.finalize()
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(ncol) * T, float(n_failed)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    value = float(tot[0].item()) / (float(t_ms.item()) * 1e-3)
    # parity pass: the seed-99 sample of the whole config that lies in this rank's block, full duration
    parity = None
    if not args.no_parity:
        sample = np.sort(np.random.default_rng(99).choice(c.n_total, args.parity_sample, replace=False)).astype(np.int64)
        mine = sample[(sample >= col0) & (sample < col1)]
        if len(mine):
            parity = parity_block(c, mine, T, min(chunk, 240), local_rank, host_cores() // max(1, world), args.mode,
                                  f"seed-99 sample of config {num} ({args.parity_sample} of {c.n_total} columns) in this rank's block, full duration, "
                                  "in an engine of their own (a column's result does not depend on its batch: tests/test_gpu_scale_parity.py)")
        if dist is not None:
            gathered = [None] * world
            dist.all_gather_object(gathered, parity)
            if rank == 0:
                parts = [p for p in gathered if p]
                parity = {k: (sum(p[k] for p in parts) if isinstance(parts[0][k], int) and k not in ("steps_per_column", "oracle_cores") else parts[0][k]) for k in parts[0]}
                parity["worst_rel"] = max(p["worst_rel"] for p in parts)
    if rank == 0:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        Bunit = 88 * 4.07 + 428
        achieved = value / world * Bunit / 1e9
        emit({"metric": "column-timesteps/sec", "value": value, "unit": "column-timesteps/s", "n_gpus": world, "steps": int(T), "warmup": 0,
              "ms_per_step": float(t_ms.item()) / T, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": {"workload": args.workload, "columns_total": int(n_total), "columns_per_gpu": int(per), "forcing_steps": int(T), "chunk_hours": chunk,
                         "what": f"BASELINE config {num} at its stated size over its full duration from the initial state, one call per chunk of forcing "
                                 "(generated on the device and resident before the call is timed), no output capture in the timed pass",
                         "l2": "inputs larger than L2: a chunk of forcing is > 10 GB per rank"},
              "clocks": sampler.summary(), "gpu_launches": int(launches), "failed_columns": int(tot[1].item()),
              "max_abs_global_mass_balance_cm": worst_gmb, "parity": parity, "e2e": None,
              "roofline": {"bound": "hbm", "kernel": "window scheduler (six stage kernels)", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                           "traffic": None, "algorithmic_bytes_per_unit": Bunit,
                           "note": "whole-path figure: B = 88 * 4.07 + 428 bytes per column-timestep (SURVEY.md 8d at the bench workload's mean front count) "
                                   "x column-timesteps / device time; per-stage figures: the default workload's line"},
              "target": {"colsteps_per_s_per_gpu": 1.25e9, "frac_of_target": value / world / 1.25e9}})
    if dist is not None:
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config5-shard", choices=["config5-shard", "config3", "config4", "config5"],
                    help="config5-shard: one GPU's share of BASELINE config 5, K steps of 48 h after a spin-up (the headline line); "
                         "config3/4/5: the BASELINE config at its stated size over its full duration, with the seed-99 parity sample in the job")
    ap.add_argument("--columns", type=int, default=1_250_000, help="columns per GPU (config5-shard)")
    ap.add_argument("--columns-total", type=int, default=0, help="config3/4/5: total columns (0: the config's stated size)")
    ap.add_argument("--hours", type=int, default=0, help="config3/4/5: forcing steps (0: the config's full duration)")
    ap.add_argument("--chunk-hours", type=int, default=1460, help="config3/4/5 and the full-year extra: forcing hours per call")
    ap.add_argument("--parity-sample", type=int, default=10_000)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--hours-per-step", type=int, default=48)
    ap.add_argument("--spinup-hours", type=int, default=720)
    ap.add_argument("--e2e-hours", type=int, default=384)
    ap.add_argument("--e2e-chunks", type=int, default=2, help="chunks of the streamed host-buffer call")
    ap.add_argument("--cpu-cols-per-core", type=int, default=2048)
    ap.add_argument("--cpu-hours", type=int, default=2880)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip value_by_call_hours, the full-year run, the stage timing window and the parity block")
    ap.add_argument("--finish-below", type=int, default=-1, help="wavefront: slots in flight below which the finishing kernel takes over (-1: engine default)")
    ap.add_argument("--mode", type=int, default=3, help="3 six stage kernels with queues between them, 1 warp-tile window kernel, 0 one kernel pair per forcing hour")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints
    # "NCCL version ..." there whatever NCCL_DEBUG_FILE says) are sent to stderr, the line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    H = args.hours_per_step
    workload = {"workload": "config5-shard", "columns_per_gpu": args.columns, "layers": 3, "giuh_ordinates": N_GIUH,
                "hours_per_step": H, "spinup_hours": args.spinup_hours, "flags": "Phillipsburg_with_reservoir (CASAM)",
                "soils": "12 parsed rows of vG_params_stat_nom_ordered.dat, i.i.d. per column and layer",
                "forcing": "synthetic hourly, P(wet)=0.03, Exp(2.5 mm/h) cap 170, diurnal/seasonal PET; counter-based, keyed by (global column, hour): "
                           "both arms and the parity block compute the same doubles (bench.forcing_np / forcing_torch)",
                "timed_region": "hours [spinup + warmup*H, + steps*H) of the simulation in ONE multi-step call, forcing resident in HBM",
                "l2": "inputs larger than L2: state + one day of forcing > 1 GB per rank, nothing is reused between hours"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        if args.workload != "config5-shard":
            emit({"impl": "reference", "unavailable": "the reference arm times the headline workload (config5-shard) only"})
            return 0
        kind = cpu_kind()
        cores = host_cores()
        cpc = args.cpu_cols_per_core
        blocks = [(H, False)] * args.warmup + [(H, True)] * args.steps
        per_block = run_cpu(kind, cores, cpc, args.spinup_hours, blocks)
        total_s = float(per_block.sum())
        colsteps = cores * cpc * H * args.steps
        val = colsteps / total_s
        line = {"impl": "reference", "metric": "column-timesteps/sec", "value": val, "unit": "column-timesteps/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload,
                "cpu_baseline": {"value": val, "unit": "column-timesteps/s", "cores": cores, "kind": kind,
                                 "sample": f"global columns 0..{cores * cpc - 1} of the workload ({cores} processes x {cpc} columns), the same hours as the B200 arm: "
                                           f"{args.spinup_hours} h spin-up, {args.warmup} warm-up and {args.steps} timed steps of {H} h, the same forcing values; "
                                           "SetValue x2 -> Update -> GetValue(total_discharge) per column-step, no file I/O"},
                "e2e": {"value": val, "unit": "column-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        line["cpu_baseline"]["standalone_executable"] = standalone_exe_entry(cores)
        emit(line)
        return 0

    import torch
    import lgar_c_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU path; use --impl reference for the CPU arm)")
    if args.workload != "config5-shard":
        return run_config_workload(args, emit, rank, world, local_rank)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    ncol = args.columns
    col0 = rank * ncol
    tmp = Path(tempfile.mkdtemp(prefix="lgar_bench_"))
    cfg = write_workload_config(tmp)
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol, local_rank, column_soils(col0, ncol), column_layers(col0, ncol))
    b.set_window_mode(args.mode)
    if args.finish_below >= 0:
        b.set_wave_params(0, 0, args.finish_below)
    ext = torch.cuda.ExternalStream(b.get_stream(), device=dev)
    dids = torch.arange(col0, col0 + ncol, dtype=torch.int64, device=dev)
    if args.steps * H * ncol * 16 > 60e9:
        raise SystemExit("bench.py: steps*hours_per_step*columns needs more than 60 GB of forcing; lower --steps or --columns")

    # spin-up and warm-up: one call per step of H hours, forcing generated per call
    hour = 0
    for _ in range(max(0, args.spinup_hours // H) + args.warmup):
        dP, dE = forcing_torch(dids, hour, H)
        torch.cuda.synchronize()
        b.update_steps_device(H, dP.data_ptr(), dE.data_ptr(), ncol)
        b.synchronize()
        hour += H
        del dP, dE
    # the forcing of the whole timed region, resident before the clock starts
    dP, dE = forcing_torch(dids, hour, H * args.steps)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    launches0 = b.get_launch_count()
    b.set_profiling(True)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    if args.mode >= 3:
        # one call for the whole timed region: columns run ahead of each other across day boundaries
        b.update_steps_device(H * args.steps, dP.data_ptr(), dE.data_ptr(), ncol)
    else:
        for k in range(args.steps):
            b.update_steps_device(H, dP[k * H].data_ptr(), dE[k * H].data_ptr(), ncol)
    ev1.record(ext)
    ev1.synchronize()
    barrier()
    sampler.stop_flag.set()
    hour += H * args.steps
    ms = ev0.elapsed_time(ev1)
    prof = b.get_profile()
    b.set_profiling(False)
    launches = b.get_launch_count() - launches0
    st = b.get_status()
    n_failed = int(np.count_nonzero(st & lgar_c_b200.STATUS_FAIL_MASK))

    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    colsteps_total = world * ncol * H * args.steps
    value = colsteps_total / (ms_max * 1e-3)

    # ---- end-to-end through the BMI verbs with host buffers -------------------------------------
    e2e = None
    if not args.no_e2e:
        # the multi-step verb with HOST buffers: forcing of `nh` hours in from pinned host memory, one output
        # series (total_discharge, ngen's usual main_output_variable) back to pinned host memory, all inside
        nh = min(args.e2e_hours, H * args.steps)
        # three pinned [hours][columns] f64 series per rank: keep all ranks of the box within half of the host memory
        nh = max(H, min(nh, int(0.5 * host_memory_bytes() / max(1, world) / (3 * 8 * ncol))))
        nh2 = nh                           # the streamed call runs over the same hours; one set of pinned buffers serves both
        hP = torch.empty((nh2, ncol), dtype=torch.float64).pin_memory()
        hE = torch.empty((nh2, ncol), dtype=torch.float64).pin_memory()
        hP.copy_(dP[:nh2])                 # the values do not matter for the timing: any hours of the workload
        hE.copy_(dE[:nh2])
        hQ = torch.empty((nh2, ncol), dtype=torch.float64).pin_memory()
        b.update_steps_host(hP.numpy()[:H], hE.numpy()[:H], {"total_discharge": hQ.numpy()[:H]})  # warm-up: allocates the staging
        barrier()
        t0 = time.perf_counter()
        b.update_steps_host(hP.numpy()[:nh], hE.numpy()[:nh], {"total_discharge": hQ.numpy()[:nh]})
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        single = world * ncol * nh / float(t_e.item())
        # the streamed verb: nh2 hours in chunks, the forcing copy of the next chunk and the output copy of the
        # previous one overlap the computation (lgarb_update_steps_host_stream)
        ch2 = (nh2 + args.e2e_chunks - 1) // max(1, args.e2e_chunks)
        # warm-up: allocates the two staging buffers per direction (chunk-sized) and the copy streams
        b.update_steps_host_stream(hP.numpy()[:ch2 + 1], hE.numpy()[:ch2 + 1], {"total_discharge": hQ.numpy()[:ch2 + 1]}, chunk_steps=ch2)
        barrier()
        t0 = time.perf_counter()
        b.update_steps_host_stream(hP.numpy(), hE.numpy(), {"total_discharge": hQ.numpy()}, chunk_steps=ch2)
        torch.cuda.synchronize()
        dt2 = time.perf_counter() - t0
        t_e2 = torch.tensor([dt2], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t_e2, op=dist.ReduceOp.MAX)
        streamed = world * ncol * nh2 / float(t_e2.item())
        e2e = {"value": max(single, streamed), "unit": "column-timesteps/s",
               "h2d_bytes_per_step": 2 * 8 * ncol * H, "d2h_bytes_per_step": 8 * ncol * H,
               "verb": "lgarb_update_steps_host_stream" if streamed >= single else "lgarb_update_steps_host",
               "single_call": {"value": single, "hours_per_call": nh, "verb": "lgarb_update_steps_host"},
               "streamed": {"value": streamed, "hours_per_call": nh2, "chunk_hours": ch2, "verb": "lgarb_update_steps_host_stream"},
               "note": "precipitation_rate + potential_evapotranspiration_rate series from pinned host memory, total_discharge series back to "
                       "pinned host memory, all copies inside the timed call; single_call copies, computes, copies back; streamed overlaps the "
                       "copies of neighbouring chunks with the computation; value = the faster of the two public verbs; bytes are per bench "
                       "step (--hours-per-step forcing hours) per rank"}
        # and the classic hour-by-hour BMI loop, for reference
        hq1 = torch.empty(ncol, dtype=torch.float64).pin_memory().numpy()
        t0 = time.perf_counter()
        nhh = min(nh, 24)
        for t in range(nhh):
            b.set_value("precipitation_rate", hP.numpy()[t])
            b.set_value("potential_evapotranspiration_rate", hE.numpy()[t])
            b.update()
            b.get_value("total_discharge", hq1)
        torch.cuda.synchronize()
        e2e["hourly_bmi_loop"] = {"value": ncol * nhh / (time.perf_counter() - t0), "unit": "column-timesteps/s (this rank)",
                                  "note": "SetValue x2, Update, GetValue(total_discharge) per forcing hour with host buffers"}
        del hP, hE, hQ

    # ---- what a caller gets for other call lengths, live stage times, a full year, parity in the job (N = 1) ----
    extras = world == 1 and not args.no_extras and args.mode >= 3
    by_call, stage_times, full_year, parity = None, None, None, None
    if extras:
        by_call = {}
        for nh in (48, 240, 960):
            if nh > H * args.steps:
                del dP, dE
                dP, dE = forcing_torch(dids, hour, nh)
            torch.cuda.synchronize()
            by_call[str(nh)] = ncol * nh / (timed_call(b, ext, nh, dP, dE, ncol) * 1e-3)
            hour += nh
        # stage times of one 240 h window, CUDA events round every stage launch (lgarb_set_stage_timing)
        b.set_stage_timing(True)
        torch.cuda.synchronize()
        ms240 = timed_call(b, ext, 240, dP, dE, ncol)
        hour += 240
        stt = b.get_stage_times()
        b.set_stage_timing(False)
        tot_ms = sum(v[0] for v in stt.values())
        stage_times = {"window_hours": 240, "window_ms": ms240,
                       "stages": {k: {"ms": v[0], "launches": v[1], "share_of_stage_time": v[0] / tot_ms if tot_ms else None} for k, v in stt.items()},
                       "note": "one extra 240 h call after the timed region with an event pair round every stage launch; CORR runs on a side stream beside "
                               "the other stages, so the stage times add up to more than the window"}
    # ---- end-of-run gather over NVLink (NCCL): the terms of lgar_global_mass_balance (reference src/lgar.cxx:1162-1216) of EVERY
    # column of every rank onto rank 0, and the all-reduced balance totals; outside the step timing, timed on its own ----
    gm = b.get_global_mass_balance()                      # [ncol][16] cm, host
    ok = (st & lgar_c_b200.STATUS_FAIL_MASK) == 0
    max_global_balance = float(np.abs(gm[ok, 15]).max()) if ok.any() else None
    gather = None
    if dist is not None:
        from lgar_c_b200.sharding import gather_summaries, reduce_totals
        summ = torch.from_numpy(gm).to(dev)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        allsum = gather_summaries(summ, dist, dst=0)
        totals = reduce_totals(torch.nan_to_num(summ).sum(dim=0), dist)
        g1.record()
        g1.synchronize()
        gms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        mgb = torch.tensor([max_global_balance or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(mgb, op=dist.ReduceOp.MAX)
        max_global_balance = float(mgb.item())
        if rank == 0:
            gather = {"ms": float(gms.item()), "columns": int(allsum.shape[0]), "values_per_column": int(allsum.shape[1]), "bytes_received": int(allsum.numel() * 8 * (world - 1) // world),
                      "totals_cm": {"precip": float(totals[1].item()), "infiltration": float(totals[2].item()), "AET": float(totals[7].item()), "discharge": float(totals[10].item())},
                      "what": "NCCL gather of the 16 terms of the global mass balance of every column of every rank onto rank 0 + all-reduce of their sums, device-timed, max over ranks"}
        del summ, allsum
    del dP, dE
    b.finalize()
    torch.cuda.empty_cache()
    if extras:
        # the full year of the config-5 shard from the initial state, forcing generated chunk by chunk (untimed) and resident before its call
        b2 = lgar_c_b200.BmiLGARBatch()
        b2.initialize(cfg, ncol, local_rank, column_soils(col0, ncol), column_layers(col0, ncol))
        b2.set_window_mode(args.mode)
        ext2 = torch.cuda.ExternalStream(b2.get_stream(), device=dev)
        yms, T, chunk_rates = 0.0, 8760, []
        for t0 in range(0, T, args.chunk_hours):
            nt = min(args.chunk_hours, T - t0)
            dP, dE = forcing_torch(dids, t0, nt)
            torch.cuda.synchronize()
            cms = timed_call(b2, ext2, nt, dP, dE, ncol)
            yms += cms
            chunk_rates.append(ncol * nt / (cms * 1e-3))
            del dP, dE
        st2 = b2.get_status()
        full_year = {"value": ncol * T / (yms * 1e-3), "unit": "column-timesteps/s", "hours": T, "chunk_hours": args.chunk_hours, "device_seconds": yms * 1e-3,
                     "by_chunk": chunk_rates,
                     "failed_columns": int(np.count_nonzero(st2 & lgar_c_b200.STATUS_FAIL_MASK)),
                     "note": "BASELINE config 5's one-year run for this GPU's 1.25 M columns from the initial state, one call per chunk of forcing"}
        by_call["8760"] = full_year["value"]
        b2.finalize()
        torch.cuda.empty_cache()
        # parity in the job: 512 columns of this shard over the hours the timed engine has simulated, against the oracle
        sys.path.insert(0, str(ROOT / "tools"))
        import parity_sampling_full as pf
        sample = col0 + np.sort(np.random.default_rng(99).choice(ncol, 512, replace=False)).astype(np.int64)
        Tp = args.spinup_hours + (args.warmup + args.steps) * H
        parity = parity_block(pf.Config(5, tmp), sample, Tp, 240, local_rank, host_cores(), args.mode,
                              f"512 columns of this shard (seed 99) over the {Tp} h the timed engine simulated (spin-up, warm-up, timed region), in an engine "
                              "of their own; the 10 000-column full-duration samples of configs 3/4/5: --workload config3|config4|config5")

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    nsteps_h = max(1, prof["steps"])
    mean_fronts = prof["front_sum"] / (nsteps_h * ncol)
    active_frac = prof["active_colsteps"] / (nsteps_h * ncol)
    L, NG = 3, N_GIUH
    B_full = 16 + 12 * L + 2 * (44 * mean_fronts + 88 + 8 * (NG + 1)) + 104      # 88*n_wf + 428
    B_nofront = B_full - 2 * 44 * mean_fronts
    ms_s, ms_a, ms_w = prof["ms_stream"], prof["ms_active"], prof["ms_window"]
    traffic_unit, share_ncu, traffic_src = committed_window_profile()
    dominant_block = None
    if ms_w > 0:  # window mode: the whole timed call is one window
        units = ncol * H
        B_unit = B_full
        dur_s = ms_w * 1e-3 / args.steps  # per bench step (H forcing hours of every column)
        dominant = "window scheduler: " + {3: "lgar_wave_kernel<FETCH..CORR> + lgar_wave_fetch_kernel + lgar_wave_finish_smem_kernel", 1: "lgar_window_kernel"}[args.mode]
        share = None
        if stage_times:
            st_ = stage_times["stages"]
            top = max(st_, key=lambda k: st_[k]["ms"])
            names = {"FETCH": "lgar_wave_fetch_kernel", "HEAD": "lgar_wave_kernel<1> (HEAD)", "FRONT": "lgar_wave_kernel<2> (FRONT)", "SEARCH": "lgar_wave_kernel<3> (SEARCH)",
                     "TAIL": "lgar_wave_kernel<4> (TAIL)", "CORR": "lgar_wave_kernel<5> (CORR)", "FINISH": "lgar_wave_finish_smem_kernel"}
            share = {names[k]: v["share_of_stage_time"] for k, v in st_.items()}
            u240 = ncol * 240
            dominant_block = {"kernel": names[top], "share_of_stage_time": st_[top]["share_of_stage_time"], "launches_per_240h_window": st_[top]["launches"],
                              "avg_launch_ms": st_[top]["ms"] / max(1, st_[top]["launches"]), "ms_per_240h_window": st_[top]["ms"],
                              "achieved_gbs_if_it_were_the_whole_path": u240 * B_unit / (st_[top]["ms"] * 1e-3) / 1e9,
                              "note": "the kernel with the largest share of the stage time of a 240 h window, timed live with CUDA events on its stream"}
    else:
        dominant = "lgar_step_active_kernel" if ms_a >= ms_s else "lgar_step_stream_kernel"
        if dominant == "lgar_step_active_kernel":
            units, B_unit, dur_s = prof["active_colsteps"] / nsteps_h, B_full, ms_a * 1e-3 / nsteps_h
        else:
            units, B_unit, dur_s = ncol, B_nofront, ms_s * 1e-3 / nsteps_h
        share = {"lgar_step_stream_kernel": ms_s / (ms_s + ms_a), "lgar_step_active_kernel": ms_a / (ms_s + ms_a)}
    bytes_per_launch = units * B_unit
    achieved = bytes_per_launch / dur_s / 1e9
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic_unit * units if (traffic_unit and args.mode == 3 and ms_w > 0) else None, "traffic_source": traffic_src,
                "stage_time_share_ncu": share_ncu, "peak_source": peak_src, "algorithmic_bytes_per_unit": B_unit,
                "units_per_launch": units, "avg_launch_ms": dur_s * 1e3, "kernel_time_share": share, "dominant": dominant_block,
                "mean_fronts": mean_fronts, "active_fraction": active_frac,
                "note": "unit = one column-timestep; a launch = one bench step (--hours-per-step forcing hours of every column; the stage kernels of a "
                        "window are timed together with CUDA events on the engine stream); B = 88*mean_fronts + 428 bytes (SURVEY.md 8d); "
                        "`frac` is the whole-path figure, `dominant` the stage kernel with the largest live share; the path is latency/divergence "
                        "bound, not HBM bound: DESIGN.md section 4.4"}

    # secondary roofline: FP64 pipe peak from a DFMA microbenchmark on this GPU
    try:
        bb = lgar_c_b200.BmiLGARBatch()
        bb.initialize(cfg, 32, local_rank)
        roofline["fp64_peak_tflops_measured"] = bb.measure_fp64_peak(40.0)
        bb.finalize()
    except Exception as ex:  # diagnostics only
        roofline["fp64_peak_tflops_measured"] = None
        print(f"bench.py: DFMA microbenchmark failed: {ex}", file=sys.stderr)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        kind = cpu_kind()
        cores = host_cores()
        cpc = args.cpu_cols_per_core
        per_block = run_cpu(kind, cores, cpc, 240, [(args.cpu_hours, True)])
        cpu = {"value": cores * cpc * args.cpu_hours / float(per_block.sum()), "unit": "column-timesteps/s", "cores": cores, "kind": kind,
               "sample": f"global columns 0..{cores * cpc - 1} of the workload ({cores} processes x {cpc} columns) x {args.cpu_hours} h after a 240 h spin-up, "
                         "the workload's own forcing values; SetValue x2 -> Update -> GetValue(total_discharge) per column-step, no file I/O"}
        cpu["standalone_executable"] = standalone_exe_entry(cores)

    line = {"metric": "column-timesteps/sec", "value": value, "unit": "column-timesteps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload, "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "failed_columns": n_failed, "max_abs_global_mass_balance_cm": max_global_balance,
            "gather": gather, "value_by_call_hours": by_call, "full_year": full_year, "stage_times": stage_times, "parity": parity, "lib_sha256": lib_sha256(),
            "target": {"colsteps_per_s_per_gpu": 1.25e9, "frac_of_target": value / world / 1.25e9}}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
