"""GPU: the asynchronous whole-call entries (explain_masses(wait=False) -> sst_explain_submit_f64 / sst_explain_collect,
classify_observed(wait=False) -> nibble-packed flags) on two context slots equal the synchronous calls, and the whole
batch equals the C oracle call for call (oracle_explain_batch_keys / oracle_is_valid_batch)."""
import numpy as np
import pytest

from oracle import oracle_c as OC
from oracle import oracle_py as OP
from spectrseqtools_b200 import fragment_classification as FC
from spectrseqtools_b200 import mass_explanation as ME
from spectrseqtools_b200 import mass_table as MT
from spectrseqtools_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _table(wl):
    seq = MT.SequenceInformation(max_len=wl.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
    return MT.DynamicProgrammingTable(S.alphabet_frame(None if len(wl.alphabet) == 104 else wl.alphabet), 32, wl.ppm, 1e-3, seq)


def _keys_per_call(batch):
    recs = np.ascontiguousarray(batch.records)
    assert recs.shape[1] == 8
    keys = recs.view(np.uint64).reshape(-1)
    call = np.repeat(np.arange(len(batch)), batch.counts())
    return keys[np.lexsort((keys, call))]


@pytest.mark.parametrize("config,n_peaks", [("C2", 6000), ("C4", 20000)])
def test_async_slots_equal_sync_and_oracle(config, n_peaks):
    wl = S.make_workload(config, n_peaks)
    dp = _table(wl)
    sync = ME.explain_masses(wl.explain_mass, dp, max_modifications=wl.max_modifications, thresholds=wl.explain_thr)
    flags = FC.classify_observed(wl.observed, dp, wl.breakage).flags.copy()
    # two batches in flight: the second half of the calls on slot 1 while slot 0 holds the first half, then both again
    half = len(wl.explain_mass) // 2
    parts = [(slice(0, half), 0), (slice(half, None), 1)]
    for _round in range(2):
        pend = [(ME.explain_masses(wl.explain_mass[s], dp, max_modifications=wl.max_modifications, thresholds=wl.explain_thr[s],
                                   wait=False, slot=k),
                 FC.classify_observed(wl.observed, dp, wl.breakage, wait=False, slot=k), s) for s, k in parts]
        for pe, pc, s in pend:
            b = pe.wait()
            lo = s.start or 0
            hi = lo + len(b)
            assert np.array_equal(b.status, sync.status[lo:hi])
            assert np.array_equal(b.counts(), sync.counts()[lo:hi])
            assert np.array_equal(b.records, sync.records[sync.offsets[lo]:sync.offsets[hi]])
            assert np.array_equal(pc.flags, flags)
    # the whole batch against the C oracle, every call
    w = [m.mass for m in dp.masses]
    tab = OC.build_bit_table(w, max(w) * wl.max_seq_length, 32)
    rows = [OP.Row(m.mass, m.is_modification, m.modification_rate) for m in dp.masses]
    ind = OP.individual_budgets(rows, dp.seq.max_len)
    tg, th = ME._integerise_many(wl.explain_mass, wl.explain_thr, dp)
    counts, keys = OC.explain_batch_keys(tab, 32, w, [r.is_modification for r in rows], ind, tg, th, wl.max_modifications, True)
    assert np.array_equal(np.where(counts < 0, 0, counts), sync.counts())
    assert np.array_equal(counts < 0, (sync.status & 2) != 0)
    assert np.array_equal(_keys_per_call(sync), keys)
    vt, vh = ME._integerise_many(wl.valid_mass, wl.valid_thr, dp)
    codes = OC.is_valid_batch(tab, 32, vt, vh)
    assert np.array_equal(np.ascontiguousarray(flags.T).reshape(-1) & 3, codes)


def test_async_falls_back_when_budgets_bind():
    """A batch whose budgets can bind cannot be queued blind: collect() carries it out synchronously, same result."""
    wl = S.make_workload("C2", 3000)
    dp = _table(wl)
    want = ME.explain_masses(wl.explain_mass, dp, max_modifications=1, thresholds=wl.explain_thr)
    got = ME.explain_masses(wl.explain_mass, dp, max_modifications=1, thresholds=wl.explain_thr, wait=False).wait()
    assert np.array_equal(got.status, want.status) and np.array_equal(got.offsets, want.offsets)
    assert np.array_equal(got.records, want.records)


def test_packed_flags_odd_fragment_count():
    wl = S.make_workload("C2", 1001)
    dp = _table(wl)
    obs = wl.observed[:1001]
    want = FC.classify_observed(obs, dp, wl.breakage).flags.copy()
    got = FC.classify_observed(obs, dp, wl.breakage, wait=False).flags
    assert got.shape == want.shape and np.array_equal(got, want)


def test_non_finite_inputs_raise_like_the_reference():
    """int(round(nan)) is a ValueError, int(round(inf)) / int(np.ceil(inf)) an OverflowError upstream
    (mass_explanation.py:51-58,107-114); the library finds them in the host arrays before anything is staged."""
    wl = S.make_workload("C2", 600)
    dp = _table(wl)
    good = wl.explain_mass[:64].copy()
    for bad, exc in ((np.nan, ValueError), (np.inf, OverflowError), (-np.inf, OverflowError)):
        m = good.copy()
        m[37] = bad
        for kw in ({}, {"wait": False}):
            with pytest.raises(exc):
                r = ME.explain_masses(m, dp, max_modifications=wl.max_modifications, thresholds=wl.explain_thr[:64], **kw)
                r.wait()
        with pytest.raises(exc):
            ME.are_valid_masses(m, dp, wl.explain_thr[:64])
        with pytest.raises(exc):
            FC.classify_observed(m, dp, wl.breakage)
        with pytest.raises(exc):
            FC.classify_observed(m, dp, wl.breakage, wait=False).wait()
        with pytest.raises(exc):
            ME.explain_mass_with_table(float(bad), dp)
        with pytest.raises(exc):
            ME.is_valid_mass(float(bad), dp)
    t = wl.explain_thr[:64].copy()
    t[5] = np.inf
    with pytest.raises(OverflowError):
        ME.explain_masses(good, dp, thresholds=t)
    t[5] = np.nan  # = None: relative threshold
    a = ME.explain_masses(good, dp, max_modifications=wl.max_modifications, thresholds=t)
    t2 = [None if np.isnan(x) else float(x) for x in t]
    b = ME.explain_masses(good, dp, max_modifications=wl.max_modifications, thresholds=t2)
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.records, b.records)
    # the context still works after the refusals
    assert len(ME.explain_masses(good, dp, max_modifications=wl.max_modifications, thresholds=wl.explain_thr[:64], wait=False).wait()) == 64


def test_speculation_recovers_after_every_kind_of_refusal():
    """sst_explain_submit_f64 queues blindly and the pass checks the staged summary on the device.  A batch it refuses
    (binding budgets, compositions longer than the record width, a heavy batch) is redone synchronously by collect()
    and the following light batch is queued blindly again — every result equals the synchronous call."""
    from spectrseqtools_b200 import _cabi

    wl = S.make_workload("C2", 3000)
    dp = _table(wl)
    ctx = _cabi.context()

    def both(masses, thr, mm):
        want = ME.explain_masses(masses, dp, max_modifications=mm, thresholds=thr)
        got = ME.explain_masses(masses, dp, max_modifications=mm, thresholds=thr, wait=False).wait()
        assert np.array_equal(got.status, want.status) and np.array_equal(got.offsets, want.offsets)
        assert got.records.shape == want.records.shape and np.array_equal(got.records, want.records)
        return got

    light = (wl.explain_mass, wl.explain_thr, wl.max_modifications)
    both(*light)
    both(wl.explain_mass, wl.explain_thr, 1)          # budgets bind: refused on the device, redone
    assert both(*light).records.shape[1] == 8          # and the light batch after it
    # 9 .. 12 nucleotides: 16-byte records, refused once, then queued blindly with the wider record
    w = sorted(m.mass for m in dp.masses if m.mass > 0)[:4]
    rng = np.random.default_rng(11)
    deep = np.array([sum(rng.choice(w, size=rng.integers(9, 13))) for _ in range(40)], dtype=np.float64) * dp.precision
    thr = np.full(len(deep), 2 * dp.precision)
    for _ in range(2):
        b = both(deep, thr, wl.max_modifications)
        assert b.records.shape[1] == 16 and b.n_compositions > 0
    assert both(*light).records.shape[1] == 8          # back to 8-byte records
    # a heavy batch (wide windows a few nucleotides up: thousands of compositions per peak) goes to the level pass
    heavy_m = np.array([sum(rng.choice(w, size=5)) for _ in range(64)], dtype=np.float64) * dp.precision
    heavy_t = np.full(len(heavy_m), 4000 * dp.precision)
    both(heavy_m, heavy_t, wl.max_modifications)
    both(*light)
    assert ctx.last_pass() == 2


def test_split_records_follow_the_longest_composition():
    """Queued batches bring their records back as planes (uint32 for the first four nucleotides + one byte plane per
    further one, learned from the previous batch): equal to the synchronous result whatever the plane count, fewer bytes
    across the bus for short compositions, and a longer batch after a shorter one is refused once and then carried."""
    from spectrseqtools_b200 import _cabi

    wl = S.make_workload("C4", 20000)
    dp = _table(wl)
    ctx = _cabi.context()
    seen, prev = {}, None
    w4 = sorted(x.mass for x in dp.masses if x.mass > 0)[:4]
    rng = np.random.default_rng(3)
    six = np.array([sum(rng.choice(w4, size=6)) for _ in range(60)], dtype=np.float64) * dp.precision  # six light nucleotides
    for max_nt in (1, 1, 9, 9, 6, 6, 1, 9):
        if max_nt == 6:
            m, t = six, np.full(len(six), 2 * dp.precision)
        else:
            sel = wl.explain_nt <= max_nt
            m, t = wl.explain_mass[sel], wl.explain_thr[sel]
        want = ME.explain_masses(m, dp, max_modifications=wl.max_modifications, thresholds=t)
        pend = ME.explain_masses(m, dp, max_modifications=wl.max_modifications, thresholds=t, wait=False, copy=False)
        got = pend.wait()
        raw = got.raw_records()
        assert np.array_equal(got.status, want.status) and np.array_equal(got.offsets, want.offsets)
        assert got.records.shape == want.records.shape and np.array_equal(got.records, want.records)
        longest = int((want.records > 0).sum(axis=1).max())
        if len(raw) > 1 or raw[0].dtype == np.uint32:  # planes: they hold the longest composition
            assert 4 + (len(raw) - 1) >= longest
            if prev == max_nt:  # (the plane count of a batch is what the batch before it needed)
                seen[max_nt] = (len(raw) - 1, ctx.explain_d2h_bytes())
        prev = max_nt
    # single nucleotides need no byte plane; six-nucleotide compositions need two when they come back as planes at all (a
    # batch the scheduler hands to the level-synchronous pass is carried out synchronously, with whole records)
    assert seen[1][0] == 0 and seen[1][0] <= seen[9][0] <= 2
    assert 6 not in seen or seen[6][0] == 2


def test_split_records_materialize():
    from spectrseqtools_b200._cabi import SplitRecords

    lo = np.array([0x04030201, 0x00000009], dtype=np.uint32)
    r = SplitRecords(lo, [np.array([5, 0], dtype=np.uint8), np.array([6, 0], dtype=np.uint8)]).materialize()
    assert r.tolist() == [[1, 2, 3, 4, 5, 6, 0, 0], [9, 0, 0, 0, 0, 0, 0, 0]]
    assert SplitRecords(np.zeros(0, dtype=np.uint32), []).materialize().shape == (0, 8)
