"""GPU: K1 (table build) and K1t (row masks) through the C-ABI, against golden vectors and the oracle."""
import numpy as np
import pytest

import helpers as Hh
import kernel_model as KM
from oracle import oracle_c as OC
from spectrseqtools_b200 import _cabi
from spectrseqtools_b200 import mass_table as MT

pytestmark = pytest.mark.gpu


def test_small_tables_bit_exact():
    cases = Hh.load_json("tables_small.json")
    arrays = np.load(Hh.GOLD / "tables_small.npz")
    done = 0
    for c in cases:
        if c.get("raises") or min(c["weights"][1:]) < 32:
            continue
        got = MT.set_up_bit_table(c["weights"], c["max_mass"], c["compression"])  # 4 / 8 / 16 per cell: host re-pack
        want = arrays[c["key"]]
        assert got.dtype == want.dtype and got.shape == want.shape
        assert np.array_equal(got, want), c
        done += 1
    assert done >= 23


def test_byte_tables_bit_exact():
    """set_up_mass_table / load_dp_table(rate 1) against tables made by the reference's own function
    (mass_table.py:292-316), incl. widths where the packed table's last word is wiped by its mask."""
    cases = Hh.load_json("mass_tables_small.json")
    arrays = np.load(Hh.GOLD / "mass_tables_small.npz")
    for c in cases:
        MT.clear_table_cache()
        got = MT.set_up_mass_table(c["weights"], c["max_mass"])
        assert got.dtype == np.uint8 and got.shape == arrays[c["key"]].shape
        assert np.array_equal(got, arrays[c["key"]]), c
    assert any((c["max_mass"] + 1) % 32 == 0 for c in cases)


def test_unsupported_tables_fail_loudly():
    with pytest.raises(ValueError):
        _cabi.context().build_table([0, 40, 50], 500, 16, 2**32 - 1)  # the device layout is 32 masses per cell only
    with pytest.raises(ValueError):
        MT.set_up_bit_table([0, 20, 50], 500, 16)      # narrow cells are re-packed from a device table: weights >= 32
    with pytest.raises(ValueError):
        MT.set_up_bit_table([0, 40, 50], 500, 5)       # not a rate the reference knows either
    with pytest.raises(ValueError):
        MT.set_up_bit_table([0, 7, 50], 500, 32)       # weight < 32: in-place loop is not a closed form
    with pytest.raises(ValueError):
        _cabi.context().build_table([0, 50, 40], 500, 32, 2**64 - 1)


@pytest.mark.parametrize("name", ["acgu", "quirk_365045", "full"])
def test_big_tables_sha(name):
    doc = Hh.load_json("tables_sha.json")[name]
    MT.clear_table_cache()
    got = MT.set_up_bit_table(doc["weights"], doc["max_mass"], 32)
    assert list(got.shape) == doc["shape"]
    assert Hh.sha(got) == doc["sha256"]
    if name == "quirk_365045":
        assert (got[:, -1] == 0).all()


def test_medium_random_alphabets_against_c_oracle():
    rng = np.random.default_rng(7)
    for trial in range(6):
        k = int(rng.integers(1, 40))
        lo = int(rng.choice([32, 1000, 1024, 5000, 30000]))
        w = [0] + sorted({int(x) for x in rng.integers(lo, lo * 3, size=k)})
        mm = max(w) * int(rng.integers(3, 36)) + int(rng.integers(0, 64))
        MT.clear_table_cache()
        got = MT.set_up_bit_table(w, mm, 32)
        want = OC.build_bit_table(w, mm, 32)
        assert np.array_equal(got, want), (trial, w[:5], mm)


def test_rebuild_is_deterministic_and_timed():
    w = Hh.full_weights()
    MT.clear_table_cache()
    dev = MT.device_table(w, max(w) * 35, 32)
    a = dev.download()
    dev.rebuild()
    b = dev.download()
    assert np.array_equal(a, b)
    build_ms, transpose_ms = dev.timings()
    assert 0 < build_ms < 1000 and 0 < transpose_ms < 1000


def test_row_masks_match_table():
    rng = np.random.default_rng(11)
    w = [0] + sorted({int(x) for x in rng.integers(1100, 4000, size=9)})
    mm = max(w) * 35
    dev = MT.device_table(w, mm, 32)
    tab = dev.download()
    masks = dev.download_masks(0, dev.limit)
    H = KM.row_masks(tab)
    got = [int(m[0]) | int(m[1]) << 32 | int(m[2]) << 64 | int(m[3]) << 96 for m in masks]
    assert got == H


def test_uploaded_table_is_adopted():
    w = [0, 1200, 1750, 2100]
    tab = OC.build_bit_table(w, 2100 * 35, 32)
    dev = _cabi.context().upload_table(tab, w)
    assert np.array_equal(dev.download(), tab)
    masks = dev.download_masks(0, dev.limit)
    H = KM.row_masks(tab)
    assert [int(m[0]) | int(m[1]) << 32 | int(m[2]) << 64 | int(m[3]) << 96 for m in masks] == H


@pytest.mark.parametrize("n_rows", [9, 40, 104])
def test_fused_row_masks_equal_the_transposed_ones(n_rows, monkeypatch):
    """SST_FUSE_MASKS=1: the build writes the mass-major row masks itself (no second kernel that reads the table back).
    Same table bytes, same masks — incl. the masked last column — as the two-kernel path, for 1, 4 and 8 rows per warp."""
    rng = np.random.default_rng(n_rows)
    w = [0] + sorted({int(x) for x in rng.integers(1100, 9000, size=3 * n_rows)})[:n_rows]
    mm = max(w) * 35 + int(rng.integers(0, 31))
    out = {}
    for fuse in ("0", "1"):
        monkeypatch.setenv("SST_FUSE_MASKS", fuse)
        MT.clear_table_cache()
        dev = MT.device_table(w, mm, 32)
        out[fuse] = (dev.download(), dev.download_masks(0, dev.limit))
    MT.clear_table_cache()
    assert np.array_equal(out["0"][0], out["1"][0])
    assert np.array_equal(out["0"][1], out["1"][1])
    assert np.array_equal(out["1"][0], OC.build_bit_table(w, mm, 32))
