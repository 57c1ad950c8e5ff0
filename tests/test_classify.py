"""N2 fused fragment classification: oracle vs reference-made golden vectors (CPU), CUDA path vs both (GPU)."""
import numpy as np
import pytest

import helpers as Hh
from oracle import oracle_c as OC
from oracle import oracle_py as OP


def _cases():
    return Hh.load_json("classify.json")


def _want(c):
    B = len(c["breakage"])
    F = len(c["observed"])
    v = np.array(c["valid"], dtype=np.uint8).reshape(B, F)
    s = np.array(c["singleton"], dtype=np.uint8).reshape(B, F)
    return v | (s << 2)


@pytest.mark.parametrize("name", ["acgu", "full_5ppm"])
def test_oracle_matches_reference_classification(name):
    c = next(x for x in _cases() if x["name"] == name)
    w = c["weights"]
    table = OC.build_bit_table(w, max(w) * 35, 32)
    got = OP.classify_pairs(c["observed"], [int(k) for k in c["breakage"]], table, w, 32, 1e-3, c["tolerance"])
    assert np.array_equal(got, _want(c))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["acgu", "full", "full_5ppm"])
def test_device_classification_matches_reference(name):
    from spectrseqtools_b200 import fragment_classification as FC

    c = next(x for x in _cases() if x["name"] == name)
    w = c["weights"]
    is_mod = [False] * len(w)
    dp = Hh.small_dp_table(w, is_mod, [0.0] + [1.0] * (len(w) - 1), 40, c["tolerance"])
    breakage = {int(k): v for k, v in c["breakage"].items()}
    res = FC.classify_observed(c["observed"], dp, breakage)
    assert np.array_equal(res.flags, _want(c))
    lazy = FC.classify_observed(c["observed"], dp, breakage, wait=False)  # side stream, waited for on first access
    assert np.array_equal(lazy.flags, _want(c))
    assert res.standard_unit_mass.shape == res.flags.shape
    # the same pairs through the plain validity call
    from spectrseqtools_b200 import mass_explanation as ME

    obs = np.array(c["observed"])
    su = res.standard_unit_mass
    for b in range(len(breakage)):
        codes = ME.are_valid_masses(su[b], dp, c["tolerance"] * obs)
        assert np.array_equal(codes, res.flags[b] & 3)


@pytest.mark.gpu
def test_classify_fragments_frame_semantics():
    """Reference-shaped wrapper: columns, order, filters (fragment_classification.py:17-101)."""
    from spectrseqtools_b200 import fragment_classification as FC
    from spectrseqtools_b200 import masses as M
    from spectrseqtools_b200 import mass_table as MT
    from spectrseqtools_b200.masses import _pl as pl

    names = M.EXPLANATION_MASSES.get_column("nucleoside").to_list()
    df = M.EXPLANATION_MASSES.filter([n in ("A", "C", "G", "U") for n in names])
    w = sorted(df.get_column("tolerated_integer_masses").to_list())
    seq_w = [w[0], w[2], w[1], w[3], w[0]]
    su_total = sum(seq_w) * 1e-3
    seq = MT.SequenceInformation(max_len=5, su_mass=su_total, obs_mass=su_total + 0.912303, modification_rate=0.5)
    dp = MT.DynamicProgrammingTable(df, 32, 10e-6, 1e-3, seq)
    breakage = M.build_breakage_dict(555.1294, 455.1491)
    lab = {v: k for k, vs in breakage.items() for v in vs}
    prefix = np.cumsum(seq_w) * 1e-3
    observed = list(prefix[:-1] + lab["START_c/y"] * 1e-3) + [su_total + lab["START_END"] * 1e-3] + [123.456, 5000.0]
    frame = pl.DataFrame({"observed_mass": observed, "intensity": [1e6] * (len(observed) - 1) + [1e6]})
    out = FC.classify_fragments(frame, dp, breakage)
    with pytest.raises(NotImplementedError):  # a mass beyond the table raises, as upstream's map_elements callback does
        FC.classify_fragments(pl.DataFrame({"observed_mass": observed + [60000.0]}), dp, breakage)
    cols = out.columns
    for c in ("fragment_index", "observed_mass", "intensity", "standard_unit_mass", "breakage", "is_singleton"):
        assert c in cols
    su = out.get_column("standard_unit_mass").to_list()
    assert su == sorted(su)
    rows = list(zip(out.get_column("fragment_index").to_list(), out.get_column("breakage").to_list(), out.get_column("is_singleton").to_list()))
    # every ladder rung is found under its true breakage, the first one is a singleton, the full sequence is START_END
    for k in range(4):
        assert any(fi == k and br == "START_c/y" for fi, br, _ in rows)
    assert any(fi == 0 and br == "START_c/y" and sg for fi, br, sg in rows)
    assert any(fi == 4 and br == "START_END" for fi, br, _ in rows)
    assert all(fi not in (5, 6) for fi, _, _ in rows)  # junk and over-heavy masses are gone
    # oracle cross-check of the kept (fragment, breakage) set
    table = OC.build_bit_table([0] + w, max(w) * 35, 32)
    flags = OP.classify_pairs(observed, list(breakage), table, [0] + w, 32, 1e-3, 10e-6)
    labels = [breakage[k][0] for k in breakage]
    want = set()
    for b in range(len(breakage)):
        for f in range(len(observed)):
            if flags[b, f] & 1:
                s = observed[f] - (list(breakage)[b] * 1e-3)
                complete = "START" in labels[b] and "END" in labels[b]
                if observed[f] < 50000 and s < su_total + 1 and (s > su_total - 1 or not complete):
                    want.add((f, labels[b]))
    assert {(fi, br) for fi, br, _ in rows} == want


@pytest.mark.gpu
@pytest.mark.parametrize("tolerance", [2e-6, 10e-6, 4e-4])
def test_device_classification_edge_windows(tolerance):
    """The lean probe path of k_classify against the Python oracle on the windows its shortcuts are made for:
    one-word windows, windows whose interior spans several words of the last-row summary (> 1024 masses wide at
    4e-4), standard-unit masses below zero, around single nucleotides (singletons), sparse low masses where the two
    end words decide, the last words of the table and beyond it (code 2)."""
    from spectrseqtools_b200 import fragment_classification as FC

    w = [0, 305042, 306026, 329053, 345048]  # A/C/G/U
    dp = Hh.small_dp_table(w, [False] * 5, [0.0] + [1.0] * 4, 40, tolerance)
    table = OC.build_bit_table(w, max(w) * 35, 32)
    limit = table.shape[1] * 32
    rng = np.random.default_rng(20260118 + int(tolerance * 1e7))
    breakage = {0: ["c/y_c/y"], 375183: ["c/y_END"], 537119: ["START_c/y"], 912303: ["START_END"], -79965: ["c/y_a/w"]}
    observed = np.concatenate([
        rng.uniform(0.0, 1500.0, 40),                       # few nucleotides: sparse table, negative SU masses
        np.array(w[1:]) * 1e-3 + rng.uniform(-0.002, 0.002, 4),  # around a single nucleotide
        np.array(w[1:]) * 1e-3 + 537.119,                    # exactly one nucleotide under START_c/y
        rng.uniform(1500.0, 9000.0, 30),                    # dense part
        (limit - rng.integers(1, 4000, 12)) * 1e-3,          # the last words of the table
        (limit + rng.integers(0, 3000, 8)) * 1e-3,           # beyond it
    ])
    res = FC.classify_observed(observed, dp, breakage)
    want = OP.classify_pairs(list(observed), list(breakage), table, w, 32, 1e-3, tolerance)
    assert np.array_equal(res.flags, want)
    assert (want & 4).any() and (want & 2).any() and (want & 1).any() and (want == 0).any()
