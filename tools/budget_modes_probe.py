#!/usr/bin/env python3
"""How the enumeration behaves when the modification budgets BIND (MEMO / EXACT modes), C4 workload."""
import pathlib, sys, time
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from spectrseqtools_b200 import synthetic as S, mass_table as MT, mass_explanation as ME

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
wl = S.make_workload("C4", n)
for rate, mm in ((0.5, wl.max_modifications), (0.05, 2), (0.02, 1), (0.0, 0)):
    seq = MT.SequenceInformation(max_len=wl.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=rate)
    dp = MT.DynamicProgrammingTable(S.alphabet_frame(None), 32, wl.ppm, 1e-3, seq)
    for with_memo in (True, False):
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            b = ME.explain_masses(wl.explain_mass, dp, max_modifications=mm, thresholds=wl.explain_thr, with_memo=with_memo, copy=False)
            ts.append(time.perf_counter() - t0)
        st = dp.device_table().ctx.kernel_stats()
        print(f"rate {rate} max_mods {mm} with_memo {with_memo}: {min(ts)*1e3:.2f} ms per call, {b.n_compositions} compositions; kernels {{k: (round(v[0],3), v[1]) for k, v in st.items() if v[1]}}".replace("{{","").replace("}}",""), {k: (round(v[0], 3), v[1]) for k, v in st.items() if v[1]})
        dp.device_table().ctx.stats_reset()
