"""Probe: the two enumeration passes side by side on the C4 batch (times, phase timestamps, equality of results)."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from spectrseqtools_b200 import _cabi, mass_explanation as ME, mass_table as MT, synthetic as S
ctx = _cabi.context(0)
wl = S.make_workload("C4", 100000)
seq = MT.SequenceInformation(max_len=wl.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
dp = MT.DynamicProgrammingTable(S.alphabet_frame(None), 32, wl.ppm, 1e-3, seq)
res = {}
flush = "--warm" not in sys.argv
for mode in (2, 3):
    ctx.set_pass(mode)
    b = ME.explain_masses(wl.explain_mass, dp, max_modifications=wl.max_modifications, thresholds=wl.explain_thr)
    res[mode] = b
    dev = dp.device_table()
    ts = []
    ctx.cta_timestamps(True)
    for _ in range(10):
        if flush:
            ctx.flush_l2()
        ctx.stats_reset(); ctx.explain_run(dev, 0); ts.append(ctx.kernel_stats()["explain_pass"][0])
    ph = ctx.explain_phase_ns().astype(np.int64); ph = ph[ph > 0]
    print("pass", mode, "last", ctx.last_pass(), "comps", b.n_compositions, "ms", np.round(ts, 4), "phases us", np.diff(ph) * 1e-3)
    if mode == 2:
        c = ctx.cta_timestamps(False).astype(np.int64)
        t0 = c[:, 0].min()
        d = (c[:, :8] - t0) * 1e-3
        for k, nm in ((0, "start"), (1, "win counted"), (2, "roots written"), (6, "rounds done"), (3, "count done"), (4, "barrier"), (5, "fill done")):
            print(f"  {nm:14s} min {d[:, k].min():8.1f} p10 {np.percentile(d[:, k], 10):8.1f} p50 {np.median(d[:, k]):8.1f} p90 {np.percentile(d[:, k], 90):8.1f} max {d[:, k].max():8.1f}")
        seg = {"windows": d[:, 1] - d[:, 0], "roots": d[:, 2] - d[:, 1], "rounds": d[:, 6] - d[:, 2], "final count": d[:, 3] - d[:, 6], "fill": d[:, 5] - d[:, 4]}
        for nm, v in seg.items():
            print(f"    {nm:12s} per CTA us p10/p50/p90/max", np.percentile(v, [10, 50, 90, 100]).round(1))
        info = c[:, 7]
        peaks, items, recs = info >> 44, (info >> 22) & 0x3FFFFF, info & 0x3FFFFF
        cnt = d[:, 3] - d[:, 0]
        print("    per CTA: peaks p10/p50/p90/max", np.percentile(peaks, [10, 50, 90, 100]).astype(int), " final items", np.percentile(items, [10, 50, 90, 100]).astype(int),
              " records", np.percentile(recs, [10, 50, 90, 100]).astype(int))
        order = np.argsort(-cnt)[:12]
        print("    slowest count phases: us / peaks / items / records / rounds us / roots us")
        for i in order:
            print(f"      {cnt[i]:6.1f} {int(peaks[i]):6d} {int(items[i]):6d} {int(recs[i]):6d} {seg['rounds'][i]:6.1f} {seg['roots'][i]:6.1f}")
        print("    corr(count us, peaks / items / records):", [round(float(np.corrcoef(cnt, x)[0, 1]), 2) for x in (peaks, items, recs)])
    if mode == 3:
        c = ctx.cta_timestamps(False).astype(np.int64)
        t0 = c[:, 0].min()
        d = (c[:, :8] - t0) * 1e-3
        names = ["start", "counted", "barrier 1", "scanned", "barrier 2", "searched", "roots (last)", "fill done"]
        for k, nm in enumerate(names):
            print(f"  {nm:14s} min {d[:, k].min():8.1f} p50 {np.median(d[:, k]):8.1f} p90 {np.percentile(d[:, k], 90):8.1f} max {d[:, k].max():8.1f}")
        print("  count phase per CTA us: p10/p50/p90/max", np.percentile(d[:, 1] - d[:, 0], [10, 50, 90, 100]).round(1),
              " fill: ", np.percentile(d[:, 7] - d[:, 4], [10, 50, 90, 100]).round(1))
a, b = res[2], res[3]
print("status", np.array_equal(a.status, b.status), "offsets", np.array_equal(a.offsets, b.offsets))
ka = a.records.view(np.uint64).reshape(-1); kb = b.records.view(np.uint64).reshape(-1)
pk = np.repeat(np.arange(len(a)), a.counts())
oa = np.lexsort((ka, pk)); ob = np.lexsort((kb, pk))
print("records (sorted per peak) equal:", np.array_equal(ka[oa], kb[ob]), " same order:", np.array_equal(ka, kb))
