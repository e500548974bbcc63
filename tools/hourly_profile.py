#!/usr/bin/env python
"""Where an hour of the classic BMI loop goes (SetValue x2 -> Update -> GetValue with host buffers, 1.25 M columns of the bench
workload after a 720 h spin-up): wall-clock seconds per verb over 24 hours.  One JSON line.
usage: [LGARB_UPDATE_WAVE_MIN_COLS=100000] python tools/hourly_profile.py [--columns N] [--hours H]"""
import argparse, json, sys, tempfile, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=1_250_000)
    ap.add_argument("--hours", type=int, default=24)
    ap.add_argument("--spinup-hours", type=int, default=720)
    a = ap.parse_args()
    import torch
    import lgar_c_b200
    ncol = a.columns
    tmp = Path(tempfile.mkdtemp(prefix="lgar_hourly_"))
    cfg = bench.write_workload_config(tmp)
    b = lgar_c_b200.BmiLGARBatch()
    b.initialize(cfg, ncol, 0, bench.column_soils(0, ncol), bench.column_layers(0, ncol))
    dids = torch.arange(0, ncol, dtype=torch.int64, device="cuda")
    hour = 0
    for _ in range(a.spinup_hours // 48):
        dP, dE = bench.forcing_torch(dids, hour, 48)
        torch.cuda.synchronize()
        b.update_steps_device(48, dP.data_ptr(), dE.data_ptr(), ncol)
        b.synchronize()
        hour += 48
    dP, dE = bench.forcing_torch(dids, hour, a.hours)
    hP, hE = dP.cpu().pin_memory().numpy(), dE.cpu().pin_memory().numpy()
    q = torch.empty(ncol, dtype=torch.float64).pin_memory().numpy()
    t = {"set_value": 0.0, "update": 0.0, "get_value": 0.0}
    l0 = b.get_launch_count()
    per_hour = []
    for k in range(a.hours):
        t0 = time.perf_counter()
        b.set_value("precipitation_rate", hP[k])
        b.set_value("potential_evapotranspiration_rate", hE[k])
        t1 = time.perf_counter()
        b.update()
        t2 = time.perf_counter()
        b.get_value("total_discharge", q)
        t3 = time.perf_counter()
        t["set_value"] += t1 - t0
        t["update"] += t2 - t1
        t["get_value"] += t3 - t2
        per_hour.append(round((t2 - t1) * 1e3, 2))
    tot = sum(t.values())
    print(json.dumps({"columns": ncol, "hours": a.hours, "column_timesteps_per_s": ncol * a.hours / tot, "ms_per_hour": {k: v / a.hours * 1e3 for k, v in t.items()},
                      "update_ms_by_hour": per_hour, "launches_per_hour": (b.get_launch_count() - l0) / a.hours,
                      "active_columns_last_hour": b.get_last_active_count()}))


if __name__ == "__main__":
    main()
