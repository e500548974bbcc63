#!/usr/bin/env python
"""Stress variant of BASELINE config 3: Bushland layering and flags, soils i.i.d. from the parsed
vG_params_stat_nom_ordered.dat, but `timestep=300[sec]` (12 sub-steps per forcing hour), flux caching off, fixed
sub-step, numeric Geff (`use_closed_form_G=false`) -- every column-step runs the heavy path.  Prints one JSON line with
the device-timed engine throughput and (with --cpu) the unmodified reference on all host cores on a sample of the
same workload.   usage: python tools/bench_stress.py [--columns N] [--hours H] [--cpu]"""
import argparse, json, sys, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench


def write_stress_config(tmpdir: Path) -> Path:
    repl = {"soil_params_file": f"soil_params_file={bench.DATA / 'vG_params_stat_nom_ordered.dat'}", "max_valid_soil_types": "max_valid_soil_types=12",
            "layer_soil_type": "layer_soil_type=1,2,3", "timestep": "timestep=300[sec]", "use_closed_form_G": "use_closed_form_G=false",
            "adaptive_timestep": "adaptive_timestep=false", "allow_flux_caching": "allow_flux_caching=false"}
    lines = []
    for line in (bench.DATA / "configs" / "config_lasam_Bushland.txt").read_text().splitlines():
        key = line.split("=", 1)[0]
        if key == "forcing_file":
            continue
        lines.append(repl.get(key, line))
    p = Path(tmpdir) / "bench_config5.txt"   # the name the CPU workers of bench.py look for
    p.write_text("\n".join(lines) + "\n")
    return p


def bushland_layers(col0, ncol):
    return np.tile(np.array(bench.BU_LAYERS, dtype=np.float64), (ncol, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=100_000)
    ap.add_argument("--hours", type=int, default=12)
    ap.add_argument("--spinup-hours", type=int, default=24)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--cpu-cols-per-core", type=int, default=16)
    args = ap.parse_args()
    bench.write_workload_config = write_stress_config
    bench.column_layers = bushland_layers
    import torch
    import lgar_c_b200
    ncol, T = args.columns, args.spinup_hours + args.hours
    tmp = Path(tempfile.mkdtemp(prefix="lgar_stress_"))
    cfg = write_stress_config(tmp)
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol, 0, bench.column_soils(0, ncol), bushland_layers(0, ncol))
    P, E = bench.host_forcing(ncol, 0, T, 777)
    dP, dE = torch.from_numpy(P).cuda(), torch.from_numpy(E).cuda()
    ext = torch.cuda.ExternalStream(b.get_stream())
    torch.cuda.synchronize()
    b.update_steps_device(args.spinup_hours, dP.data_ptr(), dE.data_ptr(), ncol)
    b.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    b.update_steps_device(args.hours, dP[args.spinup_hours].data_ptr(), dE[args.spinup_hours].data_ptr(), ncol)
    ev1.record(ext)
    ev1.synchronize()
    ms = ev0.elapsed_time(ev1)
    st = b.get_status()
    line = {"metric": "column-timesteps/sec", "workload": "config3-stress (Bushland layering, timestep=300 s, no flux caching, numeric Geff)",
            "value": ncol * args.hours / (ms * 1e-3), "unit": "column-timesteps/s", "columns": ncol, "hours": args.hours, "ms": ms,
            "substeps_per_column_step": 12, "failed_columns": int(np.count_nonzero(st & lgar_c_b200.STATUS_FAIL_MASK)), "n_gpus": 1}
    if args.cpu:
        cores = bench.host_cores()
        per_block = bench.run_cpu(bench.cpu_kind(), cores, args.cpu_cols_per_core, min(args.spinup_hours, 24), [(args.hours, True)])
        line["cpu_baseline"] = {"value": cores * args.cpu_cols_per_core * args.hours / float(per_block.sum()), "unit": "column-timesteps/s",
                                "cores": cores, "kind": bench.cpu_kind(),
                                "sample": f"{cores} processes x {args.cpu_cols_per_core} columns x {args.hours} h after a {min(args.spinup_hours, 24)} h spin-up"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
