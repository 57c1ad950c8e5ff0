"""CPU: host-side logic, the polars stand-in, and the C-ABI surface (no compute calls)."""
import math
import pathlib
import re

import numpy as np
import pytest

import helpers as Hh
from spectrseqtools_b200 import _cabi, _frame
from spectrseqtools_b200 import mass_explanation as ME
from spectrseqtools_b200 import mass_table as MT
from spectrseqtools_b200 import masses as M

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_alphabet_known_answers():
    df = M.EXPLANATION_MASSES
    assert df.columns == ["monoisotopic_mass", "nucleoside", "nucleoside_list", "modification_rate", "theoretical_mz", "tolerated_integer_masses"]
    ims = df.get_column("tolerated_integer_masses").to_list()
    assert len(ims) == 104 == len(set(ims))
    assert min(ims) == 305042 and max(ims) == 633169
    assert abs(M.PHOSPHATE_LINK_MASS - 61.956344) < 1e-12
    by_name = dict(zip(df.get_column("nucleoside").to_list(), ims))
    assert (by_name["C"], by_name["U"], by_name["A"], by_name["G"]) == (305042, 306026, 329053, 345048)
    mono = dict(zip(df.get_column("nucleoside").to_list(), df.get_column("monoisotopic_mass").to_list()))
    assert mono["A"] == 267.0968  # polars-style rounding of 267.09675, not CPython's round()
    assert M.NUC_REPS["9U"] == "U" and M.NUC_REPS["06A"] == "01A"
    assert ME.MASS_NAMES[306026] == ["U"] and ME.IS_MOD[306026] is False and ME.IS_MOD[319058] is True
    assert M.build_breakage_dict(555.1294, 455.1491) == {912303: ["START_END"], 537119: ["START_c/y"], 375183: ["c/y_END"], 0: ["c/y_c/y"]}
    assert M.TOLERANCE == 1e-3 and M.COMPRESSION_RATE == 32 and M.MATCHING_THRESHOLD == 10e-6


def test_nucleotide_mass_rows():
    rows = MT.initialize_nucleotide_masses(M.EXPLANATION_MASSES)
    assert len(rows) == 105 and rows[0].mass == 0 and rows[0].names == []
    assert [r.mass for r in rows] == sorted(r.mass for r in rows)
    assert [r.names[0] for r in rows[1:4]] == ["C", "U", "8U"]
    assert not rows[1].is_modification and rows[3].is_modification
    assert rows[1] == rows[1] and rows[1] < rows[2] and rows[2] >= rows[1]


def test_table_settings_and_mask():
    s = MT.select_table_building_settings(32)
    assert s["type"] is np.uint64 and s["init"] == 0xC000000000000000
    assert s["alt_first"] == 0xAAAAAAAAAAAAAAAA and s["alt_sec"] == 0x5555555555555555
    assert MT.select_table_building_settings(4)["alt_first"] == 0xAA
    with pytest.raises(ValueError):
        MT.select_table_building_settings(5)
    assert MT._last_column_mask(633169 * 35, 32) == (0xFFFFFFFFFFFFFFFF << 24) & 0xFFFFFFFFFFFFFFFF
    assert MT._last_column_mask(365045 * 35, 32) == 0  # the whole last word is wiped
    assert MT.MAX_SEQ_LENGTH == 35 and "dp_table" in MT.TABLE_DIR


def test_frame_stand_in_covers_the_reference_test_idioms():
    pl = _frame
    seq = ("C", "U", "A", "G")
    df = pl.DataFrame(data=seq, schema=["name"])
    lookup = M.EXPLANATION_MASSES
    df = df.with_columns(pl.col("name").map_elements(
        lambda x: lookup.filter(pl.col("nucleoside") == x).get_column("monoisotopic_mass").to_list()[0] if isinstance(lookup, pl.DataFrame)
        else dict(zip(lookup.get_column("nucleoside").to_list(), lookup.get_column("monoisotopic_mass").to_list()))[x],
        return_dtype=pl.Float64).alias("mass"))
    total = round(4 * M.PHOSPHATE_LINK_MASS + df.select("mass").sum().item(), 5)
    assert total == 1285.16888
    s = pl.Series(pl.DataFrame({"a": [3, 1, 2]}).select("a")).to_list()
    assert s == [3, 1, 2]
    j = pl.DataFrame({"k": 2}).join(pl.DataFrame({"k": [1, 2, 2], "v": ["x", "y", "z"]}), on="k", how="left")
    assert j.get_column("v").to_list() == ["y", "z"]
    assert pl.DataFrame({"a": [2, 1]}).sort("a").get_column("a").to_list() == [1, 2]


def test_budget_conversion_and_modes():
    assert ME._budget_int(np.inf) == _cabi.BUDGET_INF and ME._budget_int(None) == _cabi.BUDGET_INF
    assert ME._budget_int(2) == 2 and ME._budget_int(2.5) == 3 and ME._budget_int(-1) == 0 and ME._budget_int(0.0) == 0
    weights = np.array([0, 100, 150, 200], dtype=np.int64)
    is_mod = np.array([0, 0, 1, 1], dtype=np.uint8)
    ind = np.array([0, 9, 2, 1], dtype=np.int32)
    hi = np.array([140, 299, 300, 399, 400, -5], dtype=np.int64)
    mm = np.array([5, 5, 5, 5, 5, 0], dtype=np.int64)
    # hi=299: at most 1 copy of 150 and 1 of 200 -> free; hi=300: 2x150 ok (ind 2), 1x200 ok -> free;
    # hi=400: 2x200 > ind 1 -> bound
    assert ME._modes(weights, is_mod, ind, mm, hi, True).tolist() == [0, 0, 0, 0, 2, 0]
    assert ME._modes(weights, is_mod, ind, np.array([1] * 6), hi, False).tolist() == [0, 0, 1, 1, 1, 0]


def test_vectorised_integerisation_equals_scalar_python():
    class T:
        precision = 1e-3
        tolerance = 10e-6

    rng = np.random.default_rng(3)
    masses = np.concatenate([rng.uniform(0, 23000, 20000), np.arange(0, 2000) * 1e-3 + 0.0005, [0.0, 329.05314, 1285.16888, -3.2]])
    thr = np.concatenate([rng.uniform(0, 1.2, 10000), np.full(len(masses) - 10000, np.nan)])
    t_vec, h_vec = ME._integerise_many(masses, thr, T)
    for i in range(len(masses)):
        t, h = ME._integerise(float(masses[i]), None if np.isnan(thr[i]) else float(thr[i]), T)
        assert (t, h) == (int(t_vec[i]), int(h_vec[i])), (masses[i], thr[i])
    t2, h2 = ME._integerise_many(masses, [None if np.isnan(v) else float(v) for v in thr], T)
    assert np.array_equal(t2, t_vec) and np.array_equal(h2, h_vec)


def test_convert_names_matches_reference_semantics():
    assert ME.convert_nucleotide_masses_to_names([]).explanations is None
    assert ME.convert_nucleotide_masses_to_names([[]]).explanations == set()
    got = ME.convert_nucleotide_masses_to_names([[305042, 306026, 329053, 345048], []]).explanations
    assert got == {("C", "U", "A", "G")}


def test_abi_exports_match_header():
    header = (ROOT / "include" / "sst_b200.h").read_text()
    declared = set(re.findall(r"\b(sst_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_gpu():
    from conftest import HAVE_GPU

    if HAVE_GPU:
        pytest.skip("a GPU is present")
    with pytest.raises(_cabi.DeviceUnavailable):
        _cabi.Context(0)
    with pytest.raises(_cabi.DeviceUnavailable):
        MT.set_up_bit_table([0, 305042], 305042 * 35, 32)
    with pytest.raises(_cabi.DeviceUnavailable):
        Hh.full_dp_table(4)


def test_product_never_imports_the_oracle():
    for path in (ROOT / "spectrseqtools_b200").rglob("*.py"):
        text = path.read_text()
        assert "oracle" not in text.replace("no CPU fallback", ""), path
    for path in (ROOT / "spectrseqtools_b200" / "csrc").iterdir():
        assert "oracle" not in path.read_text(), path
