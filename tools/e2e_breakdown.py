#!/usr/bin/env python3
"""Where the end-to-end step time goes (host wall clock per call of the public batch API, C4 workload)."""
import pathlib, sys, time
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from spectrseqtools_b200 import _cabi, synthetic as S, mass_table as MT, mass_explanation as ME, fragment_classification as FC

wl = S.make_workload("C4", 100_000)
seq = MT.SequenceInformation(max_len=wl.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
dp = MT.DynamicProgrammingTable(S.alphabet_frame(None), 32, wl.ppm, 1e-3, seq)
dev = dp.device_table(); ctx = dev.ctx
obs = ctx.pinned_empty(wl.observed.shape, np.float64); obs[...] = wl.observed
em = ctx.pinned_empty(wl.explain_mass.shape, np.float64); em[...] = wl.explain_mass
et = ctx.pinned_empty(wl.explain_thr.shape, np.float64); et[...] = wl.explain_thr
offsets = np.array([w * dp.precision for w in wl.breakage])
weights, is_mod, ind = ME._row_metadata(dp)
mm = np.full(len(em), wl.max_modifications, dtype=np.int32)

def t(fn, n=30):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(n): r = fn()
    return (time.perf_counter() - t0) / n * 1e6

print("classify_stage   %.1f us" % t(lambda: ctx.classify_stage(obs, offsets)))
print("classify_run     %.1f us" % t(lambda: ctx.classify_run(dev, dp.precision, dp.tolerance)))
print("classify_fetch   %.1f us" % t(lambda: ctx.classify_fetch(copy=False)))
print("classify_observed(copy=False) %.1f us" % t(lambda: FC.classify_observed(obs, dp, wl.breakage, copy=False)))
print("_row_metadata    %.1f us" % t(lambda: ME._row_metadata(dp)))
print("explain_stage_f64 %.1f us" % t(lambda: ctx.explain_stage_f64(dev, em, et, mm, ind, is_mod, dp.precision, dp.tolerance, True)))
print("explain_run      %.1f us" % t(lambda: ctx.explain_run(dev, 0)))
print("explain_fetch(copy=False) %.1f us" % t(lambda: ctx.explain_fetch(copy=False)))
print("explain_masses(copy=False) %.1f us" % t(lambda: ME.explain_masses(em, dp, max_modifications=wl.max_modifications, thresholds=et, copy=False)))
