#!/usr/bin/env python3
"""First-visit replay (MEMO mode) of ONE deep peak over a six-row alphabet: first launch vs steady state."""
import pathlib, sys, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from spectrseqtools_b200 import masses as M, mass_table as MT, mass_explanation as ME

seq = MT.SequenceInformation(max_len=16, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
keep = {"A", "C", "G", "U", "0A", "9A"}
names = M.EXPLANATION_MASSES.get_column("nucleoside").to_list()
dp = MT.DynamicProgrammingTable(M.EXPLANATION_MASSES.filter([n in keep for n in names]), 32, 10e-6, 1e-3, seq)
w = [m.mass for m in dp.masses]
ctx = dp.device_table().ctx
for counts in ((2, 2, 2, 1, 1, 0), (3, 3, 3, 2, 1, 0), (4, 4, 3, 2, 1, 1)):
    mass = sum(c * x for c, x in zip(counts, w[1:])) * 1e-3
    for k in range(3):
        ctx.stats_reset()
        t0 = time.perf_counter()
        b = ME.explain_masses([mass], dp, max_modifications=2, with_memo=True)
        dt = time.perf_counter() - t0
        ms, n = ctx.kernel_stats()["phase_a"]
        print(f"{sum(counts)} nt, call {k}: phase_a {ms:.3f} ms x{n}, whole call {dt * 1e3:.1f} ms, {b.n_compositions} compositions")
