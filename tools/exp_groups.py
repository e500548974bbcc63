#!/usr/bin/env python
"""Experiment: the bench workload cut into G column groups, one engine + host thread per group, all on one GPU.
Measures whether concurrent windows of independent column groups fill the holes of the round scheduler
(latency-bound drop-out rounds, finishing kernel).  Wall clock around all groups, after a device synchronize.

    python tools/exp_groups.py --groups 1,2,4 --stagger-ms 0,150
"""
import argparse, sys, tempfile, threading, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", default="1,2,4")
    ap.add_argument("--stagger-ms", default="0")
    ap.add_argument("--columns", type=int, default=1_250_000)
    ap.add_argument("--hours", type=int, default=240)
    ap.add_argument("--spinup-hours", type=int, default=720)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    import torch, lgar_c_b200
    dev = torch.device("cuda", 0)
    tmp = Path(tempfile.mkdtemp(prefix="lgar_groups_"))
    cfg = bench.write_workload_config(tmp)
    ncol, H = a.columns, a.hours
    gen = torch.Generator(device=dev); gen.manual_seed(777 * 1000003)
    nh = 24 + H
    u1 = torch.rand((nh, ncol), generator=gen, device=dev, dtype=torch.float64)
    u2 = torch.rand((nh, ncol), generator=gen, device=dev, dtype=torch.float64)
    dP = torch.where(u1 < 0.03, torch.clamp(-2.5 * torch.log1p(-u2), max=170.0), torch.zeros_like(u1))
    del u1, u2
    dE = torch.from_numpy(bench.pet_series(np.arange(nh))).to(dev)[:, None].expand(nh, ncol).contiguous()
    torch.cuda.synchronize()
    for G in [int(x) for x in a.groups.split(",")]:
        bounds = [(g * ncol // G) // 32 * 32 for g in range(G)] + [ncol]
        engines = []
        for g in range(G):
            c0, c1 = bounds[g], bounds[g + 1]
            b = lgar_c_b200.BmiLGARBatch()
            b.initialize(cfg, c1 - c0, 0, bench.column_soils(c0, c1 - c0), bench.column_layers(c0, c1 - c0))
            b.set_window_mode(3)
            engines.append((b, c0, c1))

        def run(b, c0, c1, t0, nt, delay=0.0):
            if delay:
                time.sleep(delay)
            b.update_steps_device(nt, dP[t0].data_ptr() + 8 * c0, dE[t0].data_ptr() + 8 * c0, ncol)
            b.synchronize()

        def run_all(t0, nt, stagger=0.0):
            th = [threading.Thread(target=run, args=(b, c0, c1, t0, nt, g * stagger)) for g, (b, c0, c1) in enumerate(engines)]
            torch.cuda.synchronize()
            t = time.perf_counter()
            for x in th:
                x.start()
            for x in th:
                x.join()
            torch.cuda.synchronize()
            return time.perf_counter() - t

        for _ in range(a.spinup_hours // 24):
            run_all(0, 24)
        for st in [float(x) for x in a.stagger_ms.split(",")]:
            best = min(run_all(24, H, st * 1e-3) for _ in range(a.reps))
            print(f"groups {G} stagger {st:.0f} ms: {ncol * H / best:.4e} column-steps/s ({best * 1e3:.1f} ms wall for {H} h)", flush=True)
        for b, _, _ in engines:
            b.finalize()
        del engines


if __name__ == "__main__":
    main()
