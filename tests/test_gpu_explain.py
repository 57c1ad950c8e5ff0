"""GPU: K2a/K2b/K3 through the C-ABI and the reference-shaped Python API, against golden vectors + oracle."""
import math

import numpy as np
import pytest

import helpers as Hh
from oracle import oracle_c as OC
from oracle import oracle_py as OP
from spectrseqtools_b200 import _cabi
from spectrseqtools_b200 import mass_explanation as ME
from spectrseqtools_b200 import masses as M

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["auto", "level"])
def enumeration_pass(request):
    """Every test of this file runs twice: with the automatic choice (the depth-first item pass, sst_enum.cuh, wherever a
    composition has at most 16 nucleotides and the batch is light) and with the level-synchronous pass forced
    (sst_explain.cuh).  The direct pass (sst_direct.cuh) has its own file, tests/test_gpu_direct.py."""
    ctx = _cabi.context()
    ctx.set_pass({"auto": 0, "level": 1}[request.param])
    yield request.param
    ctx.set_pass(0)


def _canon(sols, w):
    index = {m: i for i, m in enumerate(w)}
    return sorted(tuple(index[x] for x in s) for s in sols if s)


def test_small_alphabets_all_budget_modes():
    """1600 reference-made cases: explain (memo / no memo), validity, None vs empty set, out-of-table."""
    cases = Hh.load_json("explain_small.json.gz")
    by_table = {}
    for c in cases:
        by_table.setdefault((tuple(c["weights"]), tuple(c["is_mod"]), tuple(c["rates"]), c["max_len"], c["tolerance"]), []).append(c)
    seen_modes = set()
    n = 0
    for (w, is_mod, rates, max_len, tol), group in by_table.items():
        dp = Hh.small_dp_table(list(w), list(is_mod), list(rates), max_len, tol)
        masses = [c["mass"] for c in group]
        thresholds = [c["threshold"] for c in group]
        mm = [np.inf if c["max_modifications"] is None else c["max_modifications"] for c in group]
        valid = ME.are_valid_masses(masses, dp, thresholds)
        for memo, tag in ((True, "memo"), (False, "nomemo")):
            batch = ME.explain_masses(masses, dp, max_modifications=mm, thresholds=thresholds, with_memo=memo)
            for p, c in enumerate(group):
                gold = c[f"solutions_{tag}"]
                if gold == "NotImplementedError":
                    assert batch.out_of_table(p)
                    with pytest.raises(NotImplementedError):
                        batch.explanations(p)
                    continue
                assert not batch.out_of_table(p)
                assert batch.canonical(p) == _canon(gold, list(w)), (c, tag)
                assert batch.has_solution(p) == (len(gold) > 0)
                assert sorted(map(tuple, batch.solutions(p))) == sorted(map(tuple, gold))
                n += 1
        for p, c in enumerate(group):
            want = c["is_valid"]
            assert int(valid[p]) == (2 if want == "NotImplementedError" else int(want)), c
    assert n >= 3000


@pytest.fixture(scope="module")
def gold_full():
    return Hh.load_json("explain_full.json")


def test_reference_unit_test_inputs(gold_full):
    """The 7 oligos x 3 tolerances of the reference's tests/test_explain_masses.py, full 104-mass alphabet,
    through the reference-shaped scalar API; full sets compared by digest."""
    for c in gold_full["unit_test_cases"]:
        dp = Hh.full_dp_table(c["max_len"], tolerance=c["tolerance"], su_mass=c["mass"], obs_mass=c["mass"])
        for memo, tag in ((True, "memo"), (False, "nomemo")):
            res = ME.explain_mass_with_table(c["mass"], dp_table=dp, max_modifications=c["max_modifications"], with_memo=memo).explanations
            assert res is not None and tuple(c["seq"]) in res
            assert len(res) == c[f"n_{tag}"], (c["seq"], c["tolerance"], tag)
            assert Hh.digest(res) == c[f"digest_{tag}"]
        if len(c["seq"]) <= 4:
            rec = ME.explain_mass_with_recursion(c["mass"], dp_table=dp, max_modifications=c["max_modifications"]).explanations
            assert Hh.digest(rec) == c["digest_recursion"] and tuple(c["seq"]) in rec


def test_full_alphabet_random_differences(gold_full, enumeration_pass):
    for c in gold_full["random_cases"]:
        dp = Hh.full_dp_table(c["max_len"])
        for memo, tag in ((True, "memo"), (False, "nomemo")):
            res = ME.explain_mass_with_table(c["mass"], dp, max_modifications=c["max_modifications"], threshold=c["threshold"], with_memo=memo).explanations
            assert (None if res is None else len(res)) == c[f"n_{tag}"], c
            assert Hh.digest(res) == c[f"digest_{tag}"], c
            assert _cabi.context().last_pass() in {"auto": (1, 2), "level": (1,)}[enumeration_pass]  # the pass under test really ran


def test_full_alphabet_validity(gold_full):
    dp = Hh.full_dp_table(35)
    cases = gold_full["validity_cases"]
    got = ME.are_valid_masses([c["mass"] for c in cases], dp, None)  # relative thresholds first
    rel = [c for c in cases if c["threshold"] is None]
    got_rel = ME.are_valid_masses([c["mass"] for c in rel], dp, None)
    for c, g in zip(rel, got_rel):
        assert int(g) == (2 if c["valid"] == "NotImplementedError" else int(c["valid"])), c
    ab = [c for c in cases if c["threshold"] is not None]
    got_ab = ME.are_valid_masses([c["mass"] for c in ab], dp, [c["threshold"] for c in ab])
    for c, g in zip(ab, got_ab):
        assert int(g) == (2 if c["valid"] == "NotImplementedError" else int(c["valid"])), c
    assert len(got) == len(cases)
    # scalar API: bool or NotImplementedError
    assert ME.is_valid_mass(329.05314, dp) is True
    assert ME.is_valid_mass(100.0, dp) is False
    with pytest.raises(NotImplementedError):
        ME.is_valid_mass(dp.device_table().limit * 1e-3, dp, 0.0)


def test_edge_cases():
    dp = Hh.full_dp_table(35)
    empty = ME.explain_masses([], dp)
    assert len(empty) == 0 and empty.n_compositions == 0
    assert len(ME.are_valid_masses([], dp)) == 0
    limit = dp.device_table().limit
    masses = [0.0, -5.0, 1e-4, 0.305042, 305.042, (limit - 1) * 1e-3, limit * 1e-3]
    batch = ME.explain_masses(masses, dp, thresholds=[0.0, 0.0, 0.001, 0.0, 0.0, 0.0, 0.0])
    assert batch.explanations(0).explanations == set()          # only the empty composition
    assert batch.explanations(1).explanations is None           # negative window
    assert batch.explanations(2).explanations == set()          # 0 inside the window
    assert batch.explanations(3).explanations is None           # 305 mDa: nothing
    assert batch.explanations(4).explanations == {("C",)}
    assert batch.explanations(5).explanations is None           # masked tail of the last word
    assert batch.out_of_table(6)
    with pytest.raises(NotImplementedError):
        batch.explanations(6)
    assert ME.is_valid_mass(633.169 * 35, dp, 0.0) is True      # 35 x heaviest = max_mass, still in the table
    ctx = dp.device_table().ctx
    ctx.set_item_limit(1000)                                  # blow-up guard: 8-nt window has far more
    try:
        with pytest.raises(MemoryError):
            ME.explain_masses([8 * 345.048], dp, thresholds=[0.01])
    finally:
        ctx.set_item_limit(0)
    with pytest.raises(ValueError):
        ME.explain_mass_with_table(305.042, dp, compression_rate=16)


def test_results_are_deterministic_and_sum_to_window():
    rng = np.random.default_rng(5)
    dp = Hh.full_dp_table(40)
    w = np.array([m.mass for m in dp.masses], dtype=np.int64)
    true = [int(w[rng.integers(1, len(w), size=int(rng.integers(1, 4)))].sum()) for _ in range(300)]
    masses = [t * 1e-3 for t in true]
    thr = [float(rng.choice([0.02, 0.06, 0.13])) for _ in true]
    a = ME.explain_masses(masses, dp, max_modifications=20, thresholds=thr)
    b = ME.explain_masses(masses, dp, max_modifications=20, thresholds=thr)
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.records, b.records)
    sums = w[a.records].sum(axis=1)
    peak_of = np.repeat(np.arange(len(masses)), a.counts())
    target = np.array([int(round(m / 1e-3, 0)) for m in masses])[peak_of]
    thr_i = np.array([int(np.ceil(t / 1e-3)) for t in thr])[peak_of]
    assert (np.abs(sums - target) <= thr_i).all()
    assert (np.diff(a.records.astype(np.int16), axis=1)[a.records[:, 1:] > 0] >= 0).all()  # rows ascending
    for p, t in enumerate(true):  # the true composition is always found
        assert a.counts()[p] >= 1


def test_against_c_oracle_on_production_like_batch():
    """A few hundred ladder-difference calls with production budgets: device counts and sets vs the C oracle."""
    rng = np.random.default_rng(9)
    dp = Hh.full_dp_table(30)
    w = [m.mass for m in dp.masses]
    tab = OC.build_bit_table(w, max(w) * 35, 32)
    rows = Hh.oracle_rows(dp)
    ind = OP.individual_budgets(rows, 30)
    is_mod = [r.is_modification for r in rows]
    masses, thr = [], []
    for _ in range(200):
        k = int(rng.integers(1, 4))
        masses.append(float(sum(w[i] for i in rng.integers(1, len(w), size=k))) * 1e-3 * (1 + rng.uniform(-2e-6, 2e-6)))
        thr.append(float(rng.choice([0.03, 0.08, 0.2])))
    batch = ME.explain_masses(masses, dp, max_modifications=15, thresholds=thr)
    for p in range(len(masses)):
        t, h = OP.integerise(masses[p], thr[p], 1e-3, 10e-6)
        r, off, _ = OC.explain(tab, 32, w, is_mod, ind, t, h, 15, True)
        want = sorted(tuple(int(x) for x in r[off[i]:off[i + 1]]) for i in range(len(off) - 1) if off[i + 1] > off[i])
        assert batch.canonical(p) == want, (p, masses[p], thr[p])


def test_memo_map_grows_on_demand():
    dp = Hh.full_dp_table(6, tolerance=10e-6)
    dev = dp.device_table()
    ctx = dev.ctx
    weights, is_mod, ind = ME._row_metadata(dp)
    target, thr = ME._integerise(1935.25876, None, dp)
    ctx.explain_stage(dev, [target], [thr], [3], [_cabi.MODE_MEMO], ind, is_mod)
    with pytest.raises(_cabi.MemoFull):
        ctx.explain_run(dev, 8, 1024)
    _, n = ctx.explain_run(dev, 8, 1 << 20)
    assert n == 792  # SURVEY Appendix C (first-visit semantics)


def test_deep_compositions_use_wide_records():
    """20-40 nucleotides per composition: records wider than 16 bytes (the run-time path-width instance of the pass),
    FREE and budgeted modes, against the C oracle."""
    w = [0, 1201, 1333, 1479]
    is_mod = [False, False, True, False]
    rates = [0.0, 1.0, 0.5, 1.0]
    max_len = 40
    dp = Hh.small_dp_table(w, is_mod, rates, max_len, 2e-6)
    tab = OC.build_bit_table(w, max(w) * 35, 32)
    rows = Hh.oracle_rows(dp)
    ind = OP.individual_budgets(rows, max_len)
    rng = np.random.default_rng(21)
    masses = [float(sum(w[i] for i in rng.integers(1, len(w), size=int(n)))) * 1e-3 for n in rng.integers(18, 34, size=24)]
    for mm, memo in ((np.inf, True), (6, True), (6, False), (25, False)):
        batch = ME.explain_masses(masses, dp, max_modifications=mm, thresholds=[0.002] * len(masses), with_memo=memo)
        assert batch.records.shape[1] >= 24
        total = 0
        for p, mass in enumerate(masses):
            t, h = OP.integerise(mass, 0.002, 1e-3, 2e-6)
            r, off, _ = OC.explain(tab, 32, w, is_mod, ind, t, h, mm, memo)
            want = sorted(tuple(int(x) for x in r[off[i]:off[i + 1]]) for i in range(len(off) - 1) if off[i + 1] > off[i])
            assert batch.canonical(p) == want, (mm, memo, p)
            total += len(want)
        assert total > 20

