"""GPU: the direct pass (sst_direct.cuh — counts looked up in the composition-count table, output-balanced fill) against
the depth-first item pass (sst_enum.cuh, which walks) record for record, and against the C oracle call for call."""
import numpy as np
import pytest

import helpers as Hh
from oracle import oracle_c as OC
from oracle import oracle_py as OP
from spectrseqtools_b200 import _cabi
from spectrseqtools_b200 import mass_explanation as ME
from spectrseqtools_b200 import mass_table as MT
from spectrseqtools_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _table(wl):
    MT.MAX_SEQ_LENGTH = wl.max_seq_length
    try:
        seq = MT.SequenceInformation(max_len=wl.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
        return MT.DynamicProgrammingTable(S.alphabet_frame(None if len(wl.alphabet) == 104 else wl.alphabet), 32, wl.ppm, 1e-3, seq)
    finally:
        MT.MAX_SEQ_LENGTH = 35


def _both(masses, dp, mm, thr):
    ctx = _cabi.context()
    try:
        ctx.set_pass(3)
        d = ME.explain_masses(masses, dp, max_modifications=mm, thresholds=thr)
        assert ctx.last_pass() == 3
        ctx.set_pass(-3)
        i = ME.explain_masses(masses, dp, max_modifications=mm, thresholds=thr)
        assert ctx.last_pass() in (1, 2)
        return d, i, ctx.last_pass()
    finally:
        ctx.set_pass(0)


@pytest.mark.parametrize("config,n_peaks", [("C2", 5000), ("C3", 5000), ("C4", 30000), ("C5", 3000)])
def test_direct_equals_item_pass_and_oracle(config, n_peaks):
    """C5 has 4-5 nt gaps at 20 ppm: thousands of compositions per call, so pieces are split through several rounds and
    cut at CTA range boundaries; C4 is the light case (one thread per peak)."""
    wl = S.make_workload(config, n_peaks)
    dp = _table(wl)
    d, i, other = _both(wl.explain_mass, dp, wl.max_modifications, wl.explain_thr)
    assert np.array_equal(d.status, i.status)
    assert np.array_equal(d.offsets, i.offsets)
    assert d.n_compositions > len(d)
    if other == 2:  # both depth-first: the very same record order
        assert np.array_equal(d.records, i.records)
    w = [m.mass for m in dp.masses]
    tab = OC.build_bit_table(w, max(w) * wl.max_seq_length, 32)
    rows = [OP.Row(m.mass, m.is_modification, m.modification_rate) for m in dp.masses]
    ind = OP.individual_budgets(rows, dp.seq.max_len)
    sel = np.arange(len(wl.explain_mass)) if config != "C5" else np.nonzero(wl.explain_nt <= 4)[0][:400]
    tg, th = ME._integerise_many(wl.explain_mass[sel], wl.explain_thr[sel], dp)
    counts, keys = OC.explain_batch_keys(tab, 32, w, [r.is_modification for r in rows], ind, tg, th, wl.max_modifications, True)
    assert np.array_equal(np.where(counts < 0, 0, counts), d.counts()[sel])
    recs = np.ascontiguousarray(d.records)
    assert recs.shape[1] == 8
    allk = recs.view(np.uint64).reshape(-1)
    got = np.concatenate([np.sort(allk[d.offsets[p]:d.offsets[p + 1]]) for p in sel]) if len(sel) else allk[:0]
    assert np.array_equal(got, keys)


def test_direct_small_alphabets_many_compositions():
    """Tiny weights: dozens of nucleotides fit a mass, counts in the thousands per window value, 16-byte records."""
    rng = np.random.default_rng(7)
    for trial in range(6):
        R = int(rng.integers(3, 9))
        w = [0] + sorted(int(x) for x in rng.choice(np.arange(40, 400), size=R, replace=False))
        dp = Hh.small_dp_table(w, [False] * (R + 1), [1.0] * (R + 1), 16, 1e-3)
        top = min(w[1] * 15, dp.device_table().C * 32 - 2)
        masses = rng.integers(1, top, size=300) * 1e-3
        thr = rng.choice([0.0, 1e-3, 4e-3], size=300)
        d, i, other = _both(masses, dp, np.inf, thr)
        assert np.array_equal(d.status, i.status) and np.array_equal(d.offsets, i.offsets)
        if other == 2:
            assert np.array_equal(d.records, i.records)
        else:
            for p in range(0, 300, 7):
                assert d.canonical(p) == i.canonical(p)
        assert d.n_compositions > 100


def test_direct_single_heavy_peak_and_empty_batch():
    wl = S.make_workload("C5", 400)
    dp = _table(wl)
    p = int(np.argmax(wl.explain_mass * (wl.explain_nt <= 5)))
    d, i, _ = _both(wl.explain_mass[p:p + 1], dp, wl.max_modifications, wl.explain_thr[p:p + 1])
    assert np.array_equal(d.offsets, i.offsets) and d.n_compositions > 50
    assert d.canonical(0) == i.canonical(0)
    e, _, _ = _both(np.zeros(0), dp, wl.max_modifications, np.zeros(0))
    assert len(e) == 0 and e.n_compositions == 0


def test_direct_declines_binding_budgets():
    wl = S.make_workload("C2", 500)
    dp = _table(wl)
    ctx = _cabi.context()
    ctx.set_pass(3)
    try:
        with pytest.raises(Exception):
            ME.explain_masses(wl.explain_mass, dp, max_modifications=1, thresholds=wl.explain_thr)
    finally:
        ctx.set_pass(0)
    ME.explain_masses(wl.explain_mass, dp, max_modifications=1, thresholds=wl.explain_thr)
    assert ctx.last_pass() in (1, 2)


def test_count_compositions_equals_the_enumeration():
    """sst_count_compositions_f64: the looked-up number of compositions of every call equals what the enumeration returns
    (budgets that cannot bind), on ladder differences and on sparse heavy ladders; the partition it feeds is balanced."""
    from spectrseqtools_b200 import sharding

    for config, n in (("C4", 20000), ("C5", 1500)):
        wl = S.make_workload(config, n)
        MT.MAX_SEQ_LENGTH = wl.max_seq_length
        try:
            seq = MT.SequenceInformation(max_len=wl.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
            dp = MT.DynamicProgrammingTable(S.alphabet_frame(None), 32, wl.ppm, 1e-3, seq)
            got = ME.count_compositions(wl.explain_mass, dp, wl.explain_thr)
            want = ME.explain_masses(wl.explain_mass, dp, max_modifications=wl.max_modifications, thresholds=wl.explain_thr).counts()
            known = got != np.uint64(2**64 - 1)
            assert known.mean() > 0.99
            assert np.array_equal(got[known].astype(np.int64), want[known])
            cuts = sharding.partition_contiguous(wl.explain_mass, wl.explain_thr, 8, dp, counts=got)
            per = np.array([want[cuts[r]:cuts[r + 1]].sum() for r in range(8)], dtype=np.float64)
            assert per.max() <= 1.15 * per.mean() + want.max()
        finally:
            MT.MAX_SEQ_LENGTH = 35
