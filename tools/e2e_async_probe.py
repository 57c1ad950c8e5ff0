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

WHAT = {"classify": True, "explain": True}


class _Done:
    def wait(self):
        return None


def submit(slot):
    t0 = time.perf_counter()
    v = FC.classify_observed(obs, dp, wl.breakage, copy=False, wait=False, slot=slot) if WHAT["classify"] else _Done()
    t1 = time.perf_counter()
    if not WHAT["explain"]:
        T["submit_classify"] += t1 - t0
        return v, _Done()
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

for depth, what in ((1, "ce"), (2, "ce"), (3, "ce"), (4, "ce"), (5, "ce"), (6, "ce"), (8, "ce"), (2, "e"), (3, "e"), (5, "e"), (2, "c"), (3, "c")):
    WHAT["classify"], WHAT["explain"] = "c" in what, "e" in what
    loop(10, depth)
    for k in T: T[k] = 0.0
    n = 200
    t0 = time.perf_counter(); loop(n, depth); dt = time.perf_counter() - t0
    print(f"depth {depth} {what}: step {dt / n * 1e6:.1f} us;  " + "  ".join(f"{k} {v / n * 1e6:.1f}" for k, v in T.items()))
# device timeline of the enumeration pass in the pipelined loop (CTA 0's %globaltimer at start / end of every launch)
from spectrseqtools_b200 import _cabi
for depth, what in ((3, "e"), (3, "ce")):
    WHAT["classify"], WHAT["explain"] = "c" in what, "e" in what
    loop(10, depth)
    stamps = []
    for k in range(depth):
        _cabi.context(ctx.device, k).trace_ms(True)
    trace, host = [], []
    pend = [submit(k) for k in range(depth - 1)]
    n = 60
    t_host = []
    for i in range(n):
        if i + depth - 1 < n:
            pend.append(submit((i + depth - 1) % depth))
        finish(pend.pop(0))
        t_host.append(time.perf_counter())
        trace.append(_cabi.context(ctx.device, i % depth).trace_ms(True)[:5].astype(np.float64) * 1e3)
        ph = _cabi.context(ctx.device, i % depth).explain_phase_ns().astype(np.int64)
        ph = ph[ph > 0]
        stamps.append((ph[0], ph[-1]))
    st = np.array(stamps, dtype=np.int64)
    dur = (st[:, 1] - st[:, 0]) * 1e-3
    gap = (st[1:, 0] - st[:-1, 1]) * 1e-3
    per = np.diff(st[:, 0]) * 1e-3
    print(f"timeline depth {depth} {what}: kernel us p50 {np.median(dur):.1f} max {dur.max():.1f}; start-to-start p50 {np.median(per):.1f}; gap after a kernel p50 {np.median(gap):.1f} min {gap.min():.1f} max {gap.max():.1f}")
    tr = np.array(trace)[10:]
    seg = np.diff(tr, axis=1)
    print("   per batch us p50: H2D", np.median(seg[:, 0]).round(1), " stage", np.median(seg[:, 1]).round(1), " pass (with waiting)", np.median(seg[:, 2]).round(1),
          " D2H", np.median(seg[:, 3]).round(1), " whole chain", np.median(tr[:, 4] - tr[:, 0]).round(1), " period", np.median(np.diff(tr[:, 4])).round(1))
    for r in tr[20:26]:
        print("     ", (r - tr[20, 0]).round(1))
    print("   host collect-to-collect us p50", round(float(np.median(np.diff(t_host))) * 1e6, 1))
WHAT["classify"] = WHAT["explain"] = True
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
