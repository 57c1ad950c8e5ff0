#!/usr/bin/env python3
"""Phase timestamps of the enumeration pass on the C4 batch tiled N times (throughput regime)."""
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
for _ in range(3):
    ctx.explain_run(dev, 0)
ctx.stats_reset()
for _ in range(5):
    ctx.flush_l2()
    r, c = ctx.explain_run(dev, 0)
st = ctx.kernel_stats()
ph = ctx.explain_phase_ns().astype(np.int64); ph = ph[ph > 0]
print(f"x{f}: {len(mass)} calls, {r} roots, {c} compositions, pass {st['explain_pass'][0] / 5:.3f} ms")
print("phases (us):", [round(float(x) * 1e-3, 1) for x in np.diff(ph)])
