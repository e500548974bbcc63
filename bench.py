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
# DRAM bytes per column-timestep of the default scheduler: ncu over every launch of one timed window (profiles/, see the
# roofline block in main)
TRAFFIC_B_PER_UNIT = 3416.4


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


def host_forcing(ncol: int, hour0: int, nhours: int, seed: int):
    """numpy version of the synthetic forcing (used by the CPU arm)."""
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
    hour = 0
    if spin_h:
        P, E = host_forcing(ncol, hour, spin_h, wid)
        advance(P, E)
        hour += spin_h
    for nh, timed in blocks:
        P, E = host_forcing(ncol, hour, nh, wid)
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
    P, E = host_forcing(ncol, 0, hours, 1000 + wid)
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
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--columns", type=int, default=1_250_000, help="columns per GPU")
    ap.add_argument("--hours-per-step", type=int, default=48)
    ap.add_argument("--spinup-hours", type=int, default=720)
    ap.add_argument("--e2e-hours", type=int, default=384)
    ap.add_argument("--e2e-chunks", type=int, default=2, help="chunks of the streamed host-buffer call")
    ap.add_argument("--cpu-cols-per-core", type=int, default=2048)
    ap.add_argument("--cpu-hours", type=int, default=2880)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--finish-below", type=int, default=-1, help="wavefront: slots in flight below which the finishing kernel takes over (-1: engine default)")
    ap.add_argument("--mode", type=int, default=3, help="3 six stage kernels with queues between them, 2 lanes kernel, 1 warp-tile window kernel, 0 one kernel pair per forcing hour")
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
                "forcing": "synthetic hourly, P(wet)=0.03, Exp(2.5 mm/h) cap 170, diurnal/seasonal PET",
                "l2": "inputs larger than L2: state + one day of forcing > 1 GB per rank, nothing is reused between hours"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        kind = cpu_kind()
        cores = host_cores()
        cpc = args.cpu_cols_per_core
        blocks = [(H, False)] * args.warmup + [(H, True)] * args.steps
        per_block = run_cpu(kind, cores, cpc, min(args.spinup_hours, 240), blocks)
        total_s = float(per_block.sum())
        colsteps = cores * cpc * H * args.steps
        val = colsteps / total_s
        line = {"impl": "reference", "metric": "column-timesteps/sec", "value": val, "unit": "column-timesteps/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload,
                "cpu_baseline": {"value": val, "unit": "column-timesteps/s", "cores": cores, "kind": kind,
                                 "sample": f"{cores} processes x {cpc} columns x {H} h per step, spin-up {min(args.spinup_hours, 240)} h; "
                                           "SetValue x2 -> Update -> GetValue(total_discharge) per column-step, no file I/O"},
                "e2e": {"value": val, "unit": "column-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        line["cpu_baseline"]["standalone_executable"] = standalone_exe_entry(cores)
        emit(line)
        return 0

    import torch
    import lgar_c_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU path; use --impl reference for the CPU arm)")
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

    # forcing ring on the device: R day-blocks, each [H][ncol]
    # the timed region is ONE multi-step call (steps*H hours), so its forcing must be contiguous: the ring
    # holds at least `steps` day-blocks (bounded: 8 bytes * 2 * H * ncol per block)
    R = max(2, args.steps, min(args.warmup + args.steps, 16))
    if R * H * ncol * 16 > 60e9:
        raise SystemExit("bench.py: steps*hours_per_step*columns needs more than 60 GB of forcing; lower --steps or --columns")
    gen = torch.Generator(device=dev)
    gen.manual_seed(777 * 1000003 + rank)
    dP = torch.empty((R * H, ncol), dtype=torch.float64, device=dev)
    dE = torch.empty((R * H, ncol), dtype=torch.float64, device=dev)
    pet = torch.from_numpy(pet_series(np.arange(R * H))).to(dev)
    for blk in range(R):
        u1 = torch.rand((H, ncol), generator=gen, device=dev, dtype=torch.float64)
        u2 = torch.rand((H, ncol), generator=gen, device=dev, dtype=torch.float64)
        sl = slice(blk * H, (blk + 1) * H)
        dP[sl] = torch.where(u1 < 0.03, torch.clamp(-2.5 * torch.log1p(-u2), max=170.0), torch.zeros_like(u1))
        dE[sl] = pet[sl, None].expand(H, ncol)
        del u1, u2
    torch.cuda.synchronize()

    def run_block(blk):
        o = (blk % R) * H
        b.update_steps_device(H, dP[o].data_ptr(), dE[o].data_ptr(), ncol)

    blk = 0
    for _ in range(max(0, args.spinup_hours // H)):
        run_block(blk)
        blk += 1
    for _ in range(args.warmup):
        run_block(blk)
        blk += 1
    b.synchronize()

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
        blk += args.steps
    else:
        for _ in range(args.steps):
            run_block(blk)
            blk += 1
    ev1.record(ext)
    ev1.synchronize()
    barrier()
    sampler.stop_flag.set()
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
        nh = min(args.e2e_hours, R * H)
        # three pinned [hours][columns] f64 series per rank: keep all ranks of the box within half of the host memory
        nh = max(H, min(nh, int(0.5 * host_memory_bytes() / max(1, world) / (3 * 8 * ncol))))
        nh2 = nh                           # the streamed call runs over the same hours; one set of pinned buffers serves both
        hP = torch.empty((nh2, ncol), dtype=torch.float64).pin_memory()
        hE = torch.empty((nh2, ncol), dtype=torch.float64).pin_memory()
        hP.copy_(dP[:nh2])
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

    # ---- end-of-run gather of per-column summaries over NCCL (outside every timed region) -------
    from lgar_c_b200.sharding import gather_summaries
    summary = torch.from_numpy(b.get_global_mass_balance(0, min(ncol, 65536))[:, [3, 15]].copy()).to(dev)
    summary = gather_summaries(summary, dist, dst=0)
    max_global_balance = float(summary[:, 1].abs().nan_to_num().max().item()) if summary is not None else 0.0

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
    if ms_w > 0:  # window mode: one kernel per bench step (H forcing hours for every column)
        dominant = {3: "lgar_wave_kernel<*> + lgar_wave_fetch_kernel (6 stage kernels)", 2: "lgar_lanes_kernel", 1: "lgar_window_kernel"}[args.mode]
        units = ncol * H
        B_unit = B_full
        dur_s = ms_w * 1e-3 / args.steps  # per bench step (one day of every column); mode 3 runs all steps in one call
        share = {dominant: 1.0}
    else:
        dominant = "lgar_step_active_kernel" if ms_a >= ms_s else "lgar_step_stream_kernel"
        if dominant == "lgar_step_active_kernel":
            units, B_unit, dur_s = prof["active_colsteps"] / nsteps_h, B_full, ms_a * 1e-3 / nsteps_h
        else:
            units, B_unit, dur_s = ncol, B_nofront, ms_s * 1e-3 / nsteps_h
        share = {"lgar_step_stream_kernel": ms_s / (ms_s + ms_a), "lgar_step_active_kernel": ms_a / (ms_s + ms_a)}
    bytes_per_launch = units * B_unit
    achieved = bytes_per_launch / dur_s / 1e9
    # DRAM bytes per launch unit from the committed ncu capture of one whole timed window of this command in mode 3
    # (profiles/r1_window_launches.csv, tools/prof_window.sh: dram__bytes_read.sum + dram__bytes_write.sum over the 2 725
    # launches of a 240 h window of 1.25 M columns = 1 024.9 GB = 3 416 B per column-timestep); null for any other mode
    traffic = TRAFFIC_B_PER_UNIT * units if (args.mode == 3 and ms_w > 0) else None
    share_ncu = ({"TAIL": 0.283, "HEAD": 0.162, "SEARCH": 0.157, "FRONT": 0.121, "finish": 0.115, "CORR": 0.089, "FETCH": 0.066}
                 if (args.mode == 3 and ms_w > 0) else None)
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": "ncu launch list of one 240 h window, profiles/r1_window_launches.csv (r1_session3_ncu_summary.md)" if traffic else None,
                "stage_time_share_ncu": share_ncu, "peak_source": peak_src, "algorithmic_bytes_per_unit": B_unit,
                "units_per_launch": units, "avg_launch_ms": dur_s * 1e3, "kernel_time_share": share,
                "mean_fronts": mean_fronts, "active_fraction": active_frac,
                "note": "unit = one column-timestep; a launch = one bench step (--hours-per-step forcing hours of every column; the stage kernels of a "
                        "window are timed together with CUDA events on the engine stream); B = 88*mean_fronts + 428 bytes (SURVEY.md 8d). "
                        "Measured DRAM traffic is 4.3x the algorithmic bytes (step contexts crossing HBM at every stage boundary); the path is "
                        "latency/divergence bound, not HBM bound: DESIGN.md section 4.4"}

    # secondary roofline: FP64 pipe peak from a DFMA microbenchmark on this GPU; what the stage kernels use of
    # it is an ncu counter (sm__inst_executed_pipe_fp64, profiles/): 38 % in SEARCH, below 15 % elsewhere
    try:
        roofline["fp64_peak_tflops_measured"] = b.measure_fp64_peak(40.0)
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
               "sample": f"{cores} processes x {cpc} columns x {args.cpu_hours} h of the same workload after a 240 h spin-up; "
                         "SetValue x2 -> Update -> GetValue(total_discharge) per column-step, no file I/O"}
        cpu["standalone_executable"] = standalone_exe_entry(cores)

    line = {"metric": "column-timesteps/sec", "value": value, "unit": "column-timesteps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload, "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "failed_columns": n_failed, "max_abs_global_mass_balance_cm": max_global_balance,
            "target": {"colsteps_per_s_per_gpu": 1.25e9, "frac_of_target": value / world / 1.25e9}}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
