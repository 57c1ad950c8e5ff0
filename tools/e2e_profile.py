#!/usr/bin/env python3
"""cProfile of the end-to-end step (public batch API, C4 workload): where the host time goes."""
import cProfile, pathlib, pstats, sys, time
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from spectrseqtools_b200 import synthetic as S, mass_table as MT, mass_explanation as ME, fragment_classification as FC

wl = S.make_workload("C4", 100_000)
seq = MT.SequenceInformation(max_len=wl.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
dp = MT.DynamicProgrammingTable(S.alphabet_frame(None), 32, wl.ppm, 1e-3, seq)
dev = dp.device_table(); ctx = dev.ctx
obs = ctx.pinned_empty(wl.observed.shape, np.float64); obs[...] = wl.observed
em = ctx.pinned_empty(wl.explain_mass.shape, np.float64); em[...] = wl.explain_mass
et = ctx.pinned_empty(wl.explain_thr.shape, np.float64); et[...] = wl.explain_thr

def step():
    valid = FC.classify_observed(obs, dp, wl.breakage, copy=False, wait=False)
    batch = ME.explain_masses(em, dp, max_modifications=wl.max_modifications, thresholds=et, copy=False)
    valid.wait()
    return valid, batch

for _ in range(5): step()
t0 = time.perf_counter()
for _ in range(200): step()
print("e2e step: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
