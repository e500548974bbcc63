#!/usr/bin/env python
"""Where does a window launch spend its time?  Per-column work trace of the lanes kernel on the bench workload."""
import sys, json, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench, lgar_c_b200, tempfile
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
H = 24
tmp = Path(tempfile.mkdtemp()); cfg = bench.write_workload_config(tmp)
b = lgar_c_b200.BmiLGARBatch(); b.initialize(cfg, ncol, 0, bench.column_soils(0, ncol), bench.column_layers(0, ncol))
dev = torch.device("cuda:0"); gen = torch.Generator(device=dev); gen.manual_seed(777 * 1000003)
R = 8
dP = torch.empty((R * H, ncol), dtype=torch.float64, device=dev); dE = torch.empty_like(dP)
pet = torch.from_numpy(bench.pet_series(np.arange(R * H))).to(dev)
for blk in range(R):
    u1 = torch.rand((H, ncol), generator=gen, device=dev, dtype=torch.float64); u2 = torch.rand((H, ncol), generator=gen, device=dev, dtype=torch.float64)
    sl = slice(blk * H, (blk + 1) * H)
    dP[sl] = torch.where(u1 < 0.03, torch.clamp(-2.5 * torch.log1p(-u2), max=170.0), torch.zeros_like(u1)); dE[sl] = pet[sl, None].expand(H, ncol)
torch.cuda.synchronize()
blk = 0
for _ in range(33):
    b.update_steps_device(H, dP[(blk % R) * H].data_ptr(), dE[(blk % R) * H].data_ptr(), ncol); blk += 1
b.synchronize()
b.set_work_trace(True)
for rep in range(3):
    t0 = time.perf_counter()
    b.update_steps_device(H, dP[(blk % R) * H].data_ptr(), dE[(blk % R) * H].data_ptr(), ncol); blk += 1
    b.synchronize(); dt = time.perf_counter() - t0
    it, kc = b.get_work_trace()
    st = b.get_status()
    order = np.argsort(-(it.astype(np.int64) * 3 + kc.astype(np.int64)))[:8]
    print(f"launch {rep}: {dt*1e3:.1f} ms; search trips: mean {it.mean():.1f} p99 {np.percentile(it,99):.0f} p99.99 {np.percentile(it,99.99):.0f} max {it.max()}; "
          f"other kcycles: mean {kc.mean():.1f} p99.99 {np.percentile(kc,99.99):.0f} max {kc.max()} (= {kc.max()*1024/1.9e6:.1f} ms)")
    sp = b.get_stage_profile(); tot = sum(v["cycles"] for v in sp.values())
    for n, v in sp.items():
        print(f"   stage {n:6s}: {100*v['cycles']/tot:5.1f}% of warp time, {v['turns']} turns, {v['lanes']/max(v['turns'],1):5.2f} lanes/turn, {v['cycles']/max(v['turns'],1):9.0f} cycles/turn")
    life, start = b.get_warp_times(); m = life > 0
    end = (start[m].astype(np.int64) - start[m].min()) + life[m]
    print("   warps", int(m.sum()), "end time us: p10 %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f" % tuple(np.percentile(end, [10, 50, 90, 99, 100])))
    for c in order[:3]:
        print("   col", int(c), "trips", int(it[c]), "other_ms %.2f" % (kc[c] * 1024 / 1.9e6), "status", int(st[c]), "soils", bench.column_soils(int(c), 1)[0], "layers", bench.column_layers(int(c), 1)[0])
