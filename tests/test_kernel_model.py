"""CPU: the Python model of the DEVICE algorithms (row masks, first-visit replay, path enumeration)
against the reference-made golden vectors.  Proves the restatement the kernels implement, not the kernels."""
import helpers as Hh
import kernel_model as KM
from oracle import oracle_c as OC
from oracle import oracle_py as OP


def test_model_matches_reference_on_small_alphabets():
    cases = Hh.load_json("explain_small.json.gz")
    tables = {}
    modes = {"free": 0, "memo": 0, "exact": 0}
    for c in cases[::2]:
        w = c["weights"]
        tab = tables.setdefault(tuple(w), OC.build_bit_table(w, max(w) * 35, 32))
        rows = [OP.Row(m, im, rt) for m, im, rt in zip(w, c["is_mod"], c["rates"])]
        ind = OP.individual_budgets(rows, c["max_len"])
        target, thr = OP.integerise(c["mass"], c["threshold"], 1e-3, c["tolerance"])
        mm = KM.INF if c["max_modifications"] is None else c["max_modifications"]
        for memo, tag in ((True, "memo"), (False, "nomemo")):
            gold = c[f"solutions_{tag}"]
            if gold == "NotImplementedError":
                try:
                    KM.explain(tab, w, c["is_mod"], ind, target, thr, mm, memo)
                    raise AssertionError("expected NotImplementedError")
                except NotImplementedError:
                    continue
            sols, zero, mode = KM.explain(tab, w, c["is_mod"], ind, target, thr, mm, memo)
            modes[mode] += 1
            assert sorted(tuple(w[r] for r in s) for s in sols) == sorted(tuple(s) for s in gold if s), (c, tag)
            assert zero == any(len(s) == 0 for s in gold)
    assert min(modes.values()) > 100, modes


def test_ladder_window_closed_form_is_the_reference_loop():
    """The count -> scan -> fill formulation of the two-pointer window (sst_ladder.cuh) visits exactly the pairs of the
    sequential loop (alphabet_reduction.ladder_pairs = prediction.py:296-327), in the same order — on dense, sparse and
    degenerate sorted ladders, with repeated masses."""
    import numpy as np

    from spectrseqtools_b200 import alphabet_reduction as AR

    rng = np.random.default_rng(17)
    assert KM.ladder_pairs_closed_form([], 10.0) == [] and KM.ladder_pairs_closed_form([3.0], 10.0) == []
    for trial in range(400):
        n = int(rng.integers(2, 40))
        gaps = rng.choice([0.0, 0.5, 3.0, 9.0, 11.0, 40.0], size=n, p=[0.1, 0.2, 0.3, 0.2, 0.1, 0.1])
        su = np.cumsum(gaps).tolist()
        mw = float(rng.choice([1.0, 10.0, 25.0, 1e9]))
        assert KM.ladder_pairs_closed_form(su, mw) == AR.ladder_pairs(su, mw), (su, mw)


def test_partition_by_looked_up_counts_balances_output():
    """sharding.partition_contiguous with per-call composition counts: contiguous blocks of equal output; unknown counts
    (2**64 - 1) fall back to the host-side estimate instead of swamping the sum."""
    import numpy as np

    from spectrseqtools_b200 import sharding

    class _Row:
        def __init__(self, mass):
            self.mass = mass

    class _DP:
        precision, tolerance = 1e-3, 1e-5
        masses = [_Row(0), _Row(305042), _Row(329053)]

    rng = np.random.default_rng(2)
    masses = rng.uniform(300.0, 2000.0, size=5000)
    counts = rng.integers(0, 50, size=5000).astype(np.uint64)
    counts[rng.integers(0, 5000, size=20)] = 5000  # a few heavy calls
    cuts = sharding.partition_contiguous(masses, None, 8, _DP(), counts=counts)
    assert cuts[0] == 0 and cuts[-1] == 5000 and all(a <= b for a, b in zip(cuts, cuts[1:]))
    per = np.array([float(counts[a:b].sum() + 16 * (b - a)) for a, b in zip(cuts, cuts[1:])])
    assert per.max() <= per.mean() + 5016 + 1
    counts[7] = np.uint64(2**64 - 1)
    cuts2 = sharding.partition_contiguous(masses, None, 8, _DP(), counts=counts)
    assert cuts2[-1] == 5000 and max(b - a for a, b in zip(cuts2, cuts2[1:])) < 2500
