"""GPU: the BASELINE.json configurations at reduced peak counts against the C oracle, and
size-independent properties of the full C4 batch."""
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
    MT.MAX_SEQ_LENGTH = wl.max_seq_length
    try:
        seq = MT.SequenceInformation(max_len=wl.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
        return MT.DynamicProgrammingTable(S.alphabet_frame(None if len(wl.alphabet) == 104 else wl.alphabet), 32, wl.ppm, 1e-3, seq)
    finally:
        MT.MAX_SEQ_LENGTH = 35


@pytest.mark.parametrize("config,n_peaks,n_explain,max_nt", [("C2", 4000, 300, 3), ("C3", 4000, 250, 3), ("C4", 4000, 250, 3), ("C5", 1500, 120, 4)])
def test_config_against_c_oracle(config, n_peaks, n_explain, max_nt):
    wl = S.make_workload(config, n_peaks)
    dp = _table(wl)
    w = [m.mass for m in dp.masses]
    tab = OC.build_bit_table(w, max(w) * wl.max_seq_length, 32)
    assert np.array_equal(dp.table, tab)  # the device table of this alphabet / width, byte for byte
    rows = [OP.Row(m.mass, m.is_modification, m.modification_rate) for m in dp.masses]
    ind = OP.individual_budgets(rows, dp.seq.max_len)
    is_mod = [r.is_modification for r in rows]
    # enumeration: the whole batch on the device, a subsample through the oracle
    sel = np.nonzero(wl.explain_nt <= max_nt)[0]
    batch = ME.explain_masses(wl.explain_mass[sel], dp, max_modifications=wl.max_modifications, thresholds=wl.explain_thr[sel])
    rng = np.random.default_rng(1)
    for q in rng.choice(len(sel), size=min(n_explain, len(sel)), replace=False):
        p = sel[q]
        t, h = OP.integerise(float(wl.explain_mass[p]), float(wl.explain_thr[p]), 1e-3, wl.ppm)
        r, off, _ = OC.explain(tab, 32, w, is_mod, ind, t, h, wl.max_modifications, True)
        want = sorted(tuple(int(x) for x in r[off[i]:off[i + 1]]) for i in range(len(off) - 1) if off[i + 1] > off[i])
        assert batch.canonical(int(q)) == want, (config, int(p))
    # classification: every (peak x breakage) pair on the device, a subsample through the oracle
    res = FC.classify_observed(wl.observed, dp, wl.breakage)
    weights_b = list(wl.breakage)
    for _ in range(600):
        b, f = int(rng.integers(len(weights_b))), int(rng.integers(len(wl.observed)))
        want = OP.classify_pairs([float(wl.observed[f])], [weights_b[b]], tab, w, 32, 1e-3, wl.ppm)[0, 0]
        assert int(res.flags[b, f]) == int(want), (config, b, f)
    # the flat validity call sees the same pairs
    n_off = len(weights_b)
    codes = ME.are_valid_masses(wl.valid_mass[: 200 * n_off], dp, wl.valid_thr[: 200 * n_off]).reshape(200, n_off)
    assert np.array_equal(codes, (res.flags[:, :200] & 3).T)


def test_full_c4_batch_properties():
    """10^5 peaks: every composition lies inside its window, rows ascend, no duplicates inside a peak, runs repeat."""
    wl = S.make_workload("C4", 100_000)
    dp = _table(wl)
    w = np.array([m.mass for m in dp.masses], dtype=np.int64)
    a = ME.explain_masses(wl.explain_mass, dp, max_modifications=wl.max_modifications, thresholds=wl.explain_thr)
    b = ME.explain_masses(wl.explain_mass, dp, max_modifications=wl.max_modifications, thresholds=wl.explain_thr)
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.records, b.records) and np.array_equal(a.status, b.status)
    assert a.n_compositions > 400_000 and len(a) == len(wl.explain_mass)
    sums = w[a.records].sum(axis=1)
    peak_of = np.repeat(np.arange(len(a)), a.counts())
    target, thr = ME._integerise_many(wl.explain_mass, wl.explain_thr, dp)
    assert (np.abs(sums - target[peak_of]) <= thr[peak_of]).all()
    rec = a.records.astype(np.int16)
    nz = rec[:, 1:] > 0
    assert ((np.diff(rec, axis=1) >= 0) | ~nz).all()  # rows ascending, zero padding at the end
    assert (rec[:, 0] > 0).all()
    key = a.records.view(np.uint64).reshape(-1)
    order = np.lexsort((key, peak_of))
    same = (np.diff(key[order]) == 0) & (np.diff(peak_of[order]) == 0)
    assert not same.any()
    # every call of a true 1-3 nt difference is explained; calls that touch an "a-B" decoy rung mostly are not
    real = ~wl.explain_decoy
    assert (a.counts()[real] >= 1).all()
    assert wl.explain_decoy.sum() > 2000 and (a.counts()[wl.explain_decoy] == 0).mean() > 0.5
    valid = FC.classify_observed(wl.observed, dp, wl.breakage)
    assert valid.flags.shape == (len(wl.breakage), wl.n_peaks) and not valid.out_of_table.any()
    assert valid.valid.any(axis=0)[~wl.observed_decoy].all()  # every true peak is explainable under its true breakage
    assert wl.observed_decoy.sum() > 1000
