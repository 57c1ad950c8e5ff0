#!/usr/bin/env python3
"""The enumeration pass on the C4 batch tiled N times (throughput regime), both passes side by side."""
import pathlib, sys
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from spectrseqtools_b200 import synthetic as S, mass_table as MT, mass_explanation as ME

f = int(sys.argv[1]) if len(sys.argv) > 1 else 16
wl = S.make_workload("C4", 100_000)
seq = MT.SequenceInformation(max_len=wl.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
dp = MT.DynamicProgrammingTable(S.alphabet_frame(None), 32, wl.ppm, 1e-3, seq)
dev = dp.device_table(); ctx = dev.ctx
mass, thr = np.tile(wl.explain_mass, f), np.tile(wl.explain_thr, f)
weights, is_mod, ind = ME._row_metadata(dp)
ctx.explain_stage_f64(dev, mass, thr, wl.max_modifications, ind, is_mod, dp.precision, dp.tolerance, True)
for mode in (2, 3):
    ctx.set_pass(mode)
    for _ in range(3):
        ctx.explain_run(dev, 0)
    ctx.stats_reset()
    if mode == 3:
        ctx.cta_timestamps(True)
    for _ in range(5):
        ctx.flush_l2()
        r, c = ctx.explain_run(dev, 0)
    st = ctx.kernel_stats()
    ph = ctx.explain_phase_ns().astype(np.int64); ph = ph[ph > 0]
    print(f"pass {mode} x{f}: {len(mass)} calls, {r} roots, {c} compositions, pass {st['explain_pass'][0] / 5:.3f} ms")
    print("  phases (us):", [round(float(x) * 1e-3, 1) for x in np.diff(ph)])
    if mode == 3:
        cs = ctx.cta_timestamps(False).astype(np.int64)
        d = (cs[:, :8] - cs[:, 0].min()) * 1e-3
        for k, nm in enumerate(["start", "counted", "barrier 1", "scanned", "barrier 2", "searched", "roots (last)", "fill done"]):
            print(f"  {nm:14s} min {d[:, k].min():8.1f} p50 {np.median(d[:, k]):8.1f} p90 {np.percentile(d[:, k], 90):8.1f} max {d[:, k].max():8.1f}")
