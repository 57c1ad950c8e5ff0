"""GPU: N1 sequence-length bounds (compute_sequence_length_bound) against reference-made golden vectors."""
import numpy as np
import pytest

import helpers as Hh
from oracle import oracle_c as OC
from oracle import oracle_py as OP
from spectrseqtools_b200 import mass_table as MT

pytestmark = pytest.mark.gpu


def test_small_alphabets_match_reference_bounds():
    cases = [c for c in Hh.load_json("explain_small.json.gz") if "bound_lower" in c]
    assert len(cases) > 200
    by_table = {}
    for c in cases:
        by_table.setdefault((tuple(c["weights"]), tuple(c["is_mod"]), tuple(c["rates"]), c["max_len"], c["tolerance"]), []).append(c)
    n = differ = 0
    for (w, is_mod, rates, max_len, tol), group in by_table.items():
        dp = Hh.small_dp_table(list(w), list(is_mod), list(rates), max_len, tol)
        dp.seq.modification_rate = 0.5  # the golden vectors were made with the universal rate 0.5 (gen_golden.py)
        for c in group:
            dp.seq.su_mass = c["mass"]
            dp.seq.obs_mass = c["bound_obs_mass"]
            for d in ("lower", "upper"):
                want = c[f"bound_{d}"]
                if want == "NotImplementedError":
                    with pytest.raises(NotImplementedError):
                        MT.compute_sequence_length_bound(dp, d)
                    continue
                assert MT.compute_sequence_length_bound(dp, d) == want, (c, d)
                n += 1
            if c["bound_lower"] != c["bound_upper"]:
                differ += 1
    assert n > 400 and differ > 20
    with pytest.raises(NotImplementedError):
        MT.compute_sequence_length_bound(dp, "sideways")


def test_whole_sequence_masses_against_c_oracle():
    """Production-shaped calls: reduced alphabets, 8-25 nt sequence masses, budgets that bind."""
    rng = np.random.default_rng(11)
    full = Hh.full_weights()
    for trial in range(6):
        k = int(rng.integers(4, 9))
        w = [0] + sorted(int(x) for x in rng.choice(full[1:], size=k, replace=False))
        is_mod = [False] + [bool(rng.random() < 0.5) for _ in w[1:]]
        rates = [0.0] + [0.5 if m else 1.0 for m in is_mod[1:]]
        L = int(rng.integers(8, 26))
        seq_w = rng.choice(w[1:], size=L)
        su = float(seq_w.sum()) * 1e-3
        max_len = int(su / 1e-3 / w[1])
        dp = Hh.small_dp_table(w, is_mod, rates, max_len, 10e-6)
        dp.seq.modification_rate = 0.5
        dp.seq.su_mass = su
        dp.seq.obs_mass = su + 0.912
        tab = OC.build_bit_table(w, max(w) * 35, 32)
        rows = Hh.oracle_rows(dp)
        ind = OP.individual_budgets(rows, max_len)
        t = int(round(su / 1e-3, 0))
        thr = int(np.ceil(10e-6 * dp.seq.obs_mass / 1e-3))
        for d in ("lower", "upper"):
            want = OC.length_bound(tab, 32, w, is_mod, ind, t, thr, round(0.5 * max_len), max_len, d)
            assert MT.compute_sequence_length_bound(dp, d) == want, (trial, d, w, L)
        lo, up = MT.compute_sequence_length_bounds(dp)
        assert lo <= L <= max(up, L)
