import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from spectrseqtools_b200 import mass_table as MT, mass_explanation as ME, masses as M
seq = MT.SequenceInformation(max_len=40, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
dp = MT.DynamicProgrammingTable(M.EXPLANATION_MASSES, 32, 10e-6, 1e-3, seq)
for m in (329.05314, 1285.16888):
    ME.explain_mass_with_table(m, dp, max_modifications=20)
    t0=time.perf_counter()
    for _ in range(200): r = ME.explain_mass_with_table(m, dp, max_modifications=20)
    print(m, len(r.explanations), "%.1f us per scalar call"%((time.perf_counter()-t0)/200*1e6))
    t0=time.perf_counter()
    for _ in range(200): ok = ME.is_valid_mass(m, dp)
    print("is_valid %.1f us"%((time.perf_counter()-t0)/200*1e6))
