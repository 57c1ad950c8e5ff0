"""N3 / N4: ladder-difference pairs and the explanation-based alphabet reduction, batched vs one call at a time."""
import copy

import numpy as np
import pytest

from spectrseqtools_b200 import alphabet_reduction as AR


def test_ladder_pairs_window_semantics():
    # differences <= 10 are "one nucleotide apart"
    su = [0.0, 4.0, 9.0, 30.0, 33.0, 41.0, 100.0, 105.0]
    pairs = AR.ladder_pairs(su, 10.0)
    # every pair is within the window, start < end, and the visiting order is by start then end
    assert all(s < e and su[e] - su[s] <= 10.0 for s, e in pairs)
    assert pairs == sorted(pairs)
    assert pairs == [(0, 1), (0, 2), (1, 2), (3, 4), (4, 5), (6, 7)]
    # tail behaviour of the reference's window: once `end` sits on the last fragment only `start` moves
    assert AR.ladder_pairs([0.0, 1.0, 2.0, 3.0], 10.0) == [(0, 1), (0, 2), (0, 3), (1, 3), (2, 3)]
    assert AR.ladder_pairs([], 10.0) == [] and AR.ladder_pairs([5.0], 10.0) == []


@pytest.mark.gpu
def test_fixed_point_equals_the_scalar_loop():
    from spectrseqtools_b200 import common, fragment_classification as FC, mass_explanation as ME
    from spectrseqtools_b200 import masses as M
    from spectrseqtools_b200 import mass_table as MT

    keep = {"A", "C", "G", "U", "0A", "9A", "0C", "0U", "8U", "2C", "7G", "01G"}
    names = M.EXPLANATION_MASSES.get_column("nucleoside").to_list()
    df = M.EXPLANATION_MASSES.filter([n in keep for n in names])
    by_name = dict(zip(df.get_column("nucleoside").to_list(), df.get_column("tolerated_integer_masses").to_list()))
    seq_names = ["C", "U", "A", "0A", "G", "G", "U", "7G", "C", "A"]
    seq_w = [by_name[n] for n in seq_names]
    su_total = sum(seq_w) * 1e-3
    breakage = M.build_breakage_dict(555.1294, 455.1491)
    lab = {v: k for k, vs in breakage.items() for v in vs}
    prefix = np.cumsum(seq_w) * 1e-3
    suffix = np.cumsum(seq_w[::-1]) * 1e-3
    rng = np.random.default_rng(3)
    observed = list(prefix[:-1] + lab["START_c/y"] * 1e-3) + list(suffix[:-1] + lab["c/y_END"] * 1e-3) + [su_total + lab["START_END"] * 1e-3]
    observed = [x * (1 + rng.uniform(-2e-6, 2e-6)) for x in observed] + [777.123, 1500.5]

    def fresh_table():
        seq = MT.SequenceInformation(max_len=int(su_total / 1e-3 / min(by_name.values())), su_mass=su_total,
                                     obs_mass=su_total + lab["START_END"] * 1e-3, modification_rate=0.5)
        return MT.DynamicProgrammingTable(df, 32, 10e-6, 1e-3, seq)

    from spectrseqtools_b200.masses import _pl as pl

    dp = fresh_table()
    frags = FC.classify_fragments(pl.DataFrame({"observed_mass": observed}), dp, breakage)
    su = frags.get_column("standard_unit_mass").to_list()
    obs = frags.get_column("observed_mass").to_list()
    brk = frags.get_column("breakage").to_list()
    single = frags.get_column("is_singleton").to_list()
    assert len(su) > 15 and any(single)

    # batched
    alive, expl = AR.filter_by_explanation(su, obs, brk, single, dp, df)
    alphabet_batched = [m.names for m in dp.masses]

    # the reference's loop, one call at a time (prediction.py:170-329 transcribed onto the scalar API)
    dp2 = fresh_table()
    cur = list(range(len(su)))
    old = -1
    max_weight = max(df.get_column("monoisotopic_mass").to_list()) + M.PHOSPHATE_LINK_MASS
    while old != len(dp2.masses):
        old = len(dp2.masses)
        explanations = {}
        for tag in ("START", "END"):
            idx = [i for i in cur if tag in brk[i]]
            s_su, s_obs = [su[i] for i in idx], [obs[i] for i in idx]
            for s, e in AR.ladder_pairs(s_su, max_weight):
                diff = s_su[e] - s_su[s]
                ex = common.calculate_explanations(diff, common.calculate_error_threshold(s_obs[s], s_obs[e], dp2.tolerance), dp2)
                if ex is not None and len(ex) >= 1:
                    explanations[diff] = ex
        for i in cur:
            if single[i]:
                explanations[su[i]] = common.calculate_explanations(su[i], dp2.tolerance * obs[i], dp2)
        seen = {n for ex in explanations.values() if ex is not None for e_ in ex for n in e_}
        dp2.adapt_individual_modification_rates_by_alphabet_reduction(seen)
        cur = [i for i in cur if ME.is_valid_mass(su[i], dp2, dp2.tolerance * obs[i])]

    assert list(alive) == cur
    assert alphabet_batched == [m.names for m in dp2.masses]
    assert set(expl) == set(explanations)
    for k in expl:
        a, b = expl[k], explanations[k]
        assert (a is None) == (b is None)
        if a is not None:
            assert sorted(x.nucleosides for x in a) == sorted(x.nucleosides for x in b)
    # the true nucleosides survive, something was removed
    kept = {n for names_ in alphabet_batched for n in names_}
    assert {"A", "C", "G", "U"} <= kept and len(dp.masses) < 13
