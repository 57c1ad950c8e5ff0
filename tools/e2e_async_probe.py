#!/usr/bin/env python3
"""Host wall clock of the pipelined end-to-end step (explain_masses / classify_observed with wait=False on two context
slots, C4 workload): per-call split and a cProfile of the loop."""
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
T = {"submit_classify": 0.0, "submit_explain": 0.0, "wait_classify": 0.0, "collect_explain": 0.0}

def submit(slot):
    t0 = time.perf_counter()
    v = FC.classify_observed(obs, dp, wl.breakage, copy=False, wait=False, slot=slot)
    t1 = time.perf_counter()
    b = ME.explain_masses(em, dp, max_modifications=wl.max_modifications, thresholds=et, copy=False, wait=False, slot=slot)
    t2 = time.perf_counter()
    T["submit_classify"] += t1 - t0; T["submit_explain"] += t2 - t1
    return v, b

def finish(p):
    t0 = time.perf_counter()
    p[0].wait()
    t1 = time.perf_counter()
    r = p[1].wait()
    t2 = time.perf_counter()
    T["wait_classify"] += t1 - t0; T["collect_explain"] += t2 - t1
    return r

def loop(n, depth=2):
    pend = [submit(k) for k in range(depth - 1)]
    for i in range(n):
        if i + depth - 1 < n:
            pend.append(submit((i + depth - 1) % depth))
        finish(pend.pop(0))

for depth in (1, 2):
    loop(10, depth)
    for k in T: T[k] = 0.0
    n = 200
    t0 = time.perf_counter(); loop(n, depth); dt = time.perf_counter() - t0
    print(f"depth {depth}: step {dt / n * 1e6:.1f} us;  " + "  ".join(f"{k} {v / n * 1e6:.1f}" for k, v in T.items()))
ctx.host_profile(True)
n = 200
loop(n, 2)
ns, calls = ctx.host_profile(False)
names = {0: "submit: scans of the host arrays", 1: "submit: mode bound + cost sample", 2: "stage: reserves", 3: "stage: memset", 4: "stage: H2D copies",
         5: "stage: k_stage_f64 launch", 6: "stage: summary D2H", 7: "dfs: reserves + args", 8: "dfs: timer event", 9: "dfs: cooperative launch",
         10: "dfs: 2 event records", 11: "D2H status", 12: "D2H offsets + records", 13: "collect: stream wait", 16: "classify: wait stream2",
         17: "classify: scan", 18: "classify: 2 H2D", 19: "classify: launch", 20: "classify: D2H"}
for k in range(32):
    if calls[k]:
        print(f"  [{k:2d}] {names.get(k, ''):36s} {ns[k] / n * 1e-3:8.1f} us/step  ({int(calls[k])} visits)")
pr = cProfile.Profile(); pr.enable(); loop(200, 2); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(8)
