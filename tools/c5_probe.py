#!/usr/bin/env python3
"""C5 (combinatorial blow-up) probe: sparse 80-nt ladders, 1-5 nt gaps, 20 ppm, table to 26 kDa.
Reports compositions/s of the enumeration pass and the bytes it writes."""
import pathlib, sys, time
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from spectrseqtools_b200 import synthetic as S, mass_table as MT, mass_explanation as ME

n_peaks = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
max_nt = int(sys.argv[2]) if len(sys.argv) > 2 else 5
wl = S.make_workload("C5", n_peaks)
MT.MAX_SEQ_LENGTH = wl.max_seq_length
seq = MT.SequenceInformation(max_len=wl.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
dp = MT.DynamicProgrammingTable(S.alphabet_frame(None), 32, wl.ppm, 1e-3, seq)
dev = dp.device_table(); ctx = dev.ctx
sel = wl.explain_nt <= max_nt
em, et = wl.explain_mass[sel], wl.explain_thr[sel]
print("calls", len(em), "by nt", np.bincount(wl.explain_nt[sel]), "table", dev.R, "x", dev.C)
weights, is_mod, ind = ME._row_metadata(dp)
mm = np.full(len(em), wl.max_modifications, dtype=np.int32)
ctx.explain_stage_f64(dev, em, et, mm, ind, is_mod, dp.precision, dp.tolerance, True)
for which in (1, -3, 3):
    ctx.set_pass(which)
    t0 = time.perf_counter(); nr, nc = ctx.explain_run(dev, 0); t1 = time.perf_counter()
    print("pass choice %d (ran %d): first run (buffers grow): %.1f ms, roots %d comps %d" % (which, ctx.last_pass(), (t1 - t0) * 1e3, nr, nc))
    ctx.stats_reset()
    for _ in range(3):
        ctx.flush_l2(); ctx.timer_start(); nr, nc = ctx.explain_run(dev, 0); ms = ctx.timer_stop()
        print("  run: %.3f ms  %.3g comps/s  records %.1f MB -> %.0f GB/s written" % (ms, nc / ms * 1e3, nc * 8 / 1e6, nc * 8 / ms / 1e6))
    print(" ", ctx.kernel_stats()["explain_pass"], [round(float(x) * 1e-3, 1) for x in np.diff(ctx.explain_phase_ns().astype(np.int64)[ctx.explain_phase_ns() > 0])])
