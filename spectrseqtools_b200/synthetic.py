"""Synthetic spectra for benchmarks and scale tests (BASELINE.json configs C2-C5).

An oligo is L nucleotides drawn from the configured table rows (15 % modified, at most half).  Its 5'
ladder (prefix fragments, START tag) and 3' ladder (suffix fragments, END tag) give 2L observed peaks
with multiplicative noise.  From them come the two kinds of calls the reference pipeline makes:

* validity calls — every peak x every breakage offset, threshold tol * observed
  (fragment_classification.py:52-60 upstream);
* explanation calls — differences of adjacent kept ladder rungs (20 % of the rungs are missing, gaps of
  more than three nucleotides are skipped), threshold tol * (obs_a + obs_b) (common.py:37-40), plus one
  singleton call per 1-nt peak.

Everything is seeded: numpy.random.default_rng(20260118 + config_id).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import masses as M

C2_MODS = "0C 0U 8U 2C 2U 9A 0A 04C 03U 01A 68A 7G 01G 071C 61A 62A 10G 51C 022G 2511U".split()


@dataclass
class Workload:
    name: str
    config_id: int
    ppm: float                 # relative tolerance as a fraction (10e-6 = 10 ppm)
    alphabet: List[str]        # representative names of the table rows
    max_len: int               # seq.max_len for the table (upper bound over the oligos)
    max_seq_length: int        # MAX_SEQ_LENGTH the table needs (35 upstream, 42 for C5)
    n_peaks: int
    valid_mass: np.ndarray     # float64 SU-mass candidates, n_peaks * n_offsets
    valid_thr: np.ndarray      # float64 absolute thresholds
    explain_mass: np.ndarray   # float64 ladder differences / singleton masses
    explain_thr: np.ndarray
    explain_nt: np.ndarray     # true number of nucleotides of each difference (for reporting)
    max_modifications: int
    observed: Optional[np.ndarray] = None   # float64 observed peak masses, n_peaks (valid_* = observed x breakage)
    breakage: Optional[dict] = None         # breakage weight (integer mDa, ascending) -> labels
    observed_decoy: Optional[np.ndarray] = None   # bool, n_peaks: the peak is an "a-B" decoy (C4)
    explain_decoy: Optional[np.ndarray] = None    # bool per explanation call: one of its rungs is a decoy


def alphabet_frame(names: Optional[List[str]]):
    """EXPLANATION_MASSES restricted to the given representatives (None = full alphabet)."""
    df = M.EXPLANATION_MASSES
    if names is None:
        return df
    keep = set(names)
    col = df.get_column("nucleoside").to_list()
    missing = keep - set(col)
    if missing:
        raise ValueError(f"not representatives of the alphabet: {sorted(missing)}")
    return df.filter([n in keep for n in col])


def _rows(names: Optional[List[str]]):
    df = alphabet_frame(names)
    ims = np.array(df.get_column("tolerated_integer_masses").to_list(), dtype=np.int64)
    reps = df.get_column("nucleoside").to_list()
    order = np.argsort(ims)
    ims = ims[order]
    reps = [reps[i] for i in order]
    is_mod = np.array([r not in M.UNMODIFIED_BASES for r in reps])
    return ims, reps, is_mod


NEUTRAL_BASE_MASSES = np.array([135.05450, 111.04326, 151.04941, 112.02728])  # adenine, cytosine, guanine, uracil


def make_frame(n_oligos: int, length: int = 25, seed: int = 7, names=None, miss_p: float = 0.2, ppm: float = 10e-6):
    """A classified-fragment frame for the ladder rows (N3 / N4): both ladders of ``n_oligos`` random oligos, some rungs
    missing, sorted by standard-unit mass as prediction.py:68-72 leaves them.  -> (su, observed, breakage labels,
    is_singleton): the columns ``Predictor.filter_by_explanation`` reads."""
    rng = np.random.default_rng(seed)
    ims, reps, is_mod = _rows(names)
    plain, mods = np.nonzero(~is_mod)[0], np.nonzero(is_mod)[0]
    breakage = M.build_breakage_dict(555.1294, 455.1491)
    label_of = {v: k for k, vs in breakage.items() for v in vs}
    off5, off3 = label_of["START_c/y"] * 1e-3, label_of["c/y_END"] * 1e-3
    su_all, obs_all, brk, single = [], [], [], []
    for _ in range(n_oligos):
        pick = np.where(rng.random(length) < 0.15, rng.choice(mods, size=length), rng.choice(plain, size=length))
        w = ims[pick] * 1e-3
        for ladder, (off, tag) in enumerate(((off5, "START_c/y"), (off3, "c/y_END"))):
            su = np.cumsum(w if ladder == 0 else w[::-1])
            obs = (su + off) * (1 + rng.uniform(-ppm / 2, ppm / 2, size=length))
            kept = np.nonzero(rng.random(length) >= miss_p)[0]
            su_all.append((obs - off)[kept])
            obs_all.append(obs[kept])
            brk += [tag] * len(kept)
            single += [bool(k == 0) for k in kept]
    su, obs = np.concatenate(su_all), np.concatenate(obs_all)
    order = np.argsort(su, kind="stable")
    return su[order], obs[order], [brk[i] for i in order], np.array(single, dtype=bool)[order]


def make_workload(config: str, n_peaks: int = 100_000, seed_offset: int = 0) -> Workload:
    config = config.upper()
    miss_p, max_gap = 0.2, 3
    if config == "C2":
        cid, lens, names, ppm, msl, full_dict = 2, (20, 20), ["A", "C", "G", "U"] + C2_MODS, 5e-6, 35, False
    elif config == "C3":
        cid, lens, names, ppm, msl, full_dict = 3, (40, 40), None, 10e-6, 35, False
    elif config == "C4":
        cid, lens, names, ppm, msl, full_dict = 4, (10, 40), None, 10e-6, 35, True
    elif config == "C5":
        cid, lens, names, ppm, msl, full_dict = 5, (80, 80), None, 20e-6, 42, False
        miss_p, max_gap = 0.55, 5  # sparse ladders: 1-5 nt gaps, the combinatorial blow-up case
    else:
        raise ValueError(f"unknown workload {config!r} (C2, C3, C4, C5)")
    rng = np.random.default_rng(20260118 + cid + seed_offset)
    ims, reps, is_mod = _rows(names)
    plain = np.nonzero(~is_mod)[0]
    mods = np.nonzero(is_mod)[0]

    old = M.FULL_BREAKAGE_DICT
    try:
        M.FULL_BREAKAGE_DICT = full_dict
        breakage = M.build_breakage_dict(555.1294, 455.1491)
    finally:
        M.FULL_BREAKAGE_DICT = old
    offsets = np.array(sorted(breakage), dtype=np.int64)  # integer mDa
    label_of = {v: k for k, vs in breakage.items() for v in vs}
    off5 = label_of["START_c/y"] * 1e-3
    off3 = label_of["c/y_END"] * 1e-3
    if full_dict:  # C4: half of the ladders are a/w-type ions
        off5_alt = label_of["START_a/w"] * 1e-3
        off3_alt = label_of["a/w_END"] * 1e-3

    v_mass, v_thr, e_mass, e_thr, e_nt, all_obs, all_decoy, e_decoy = [], [], [], [], [], [], [], []
    peaks = 0
    longest = 0
    while peaks < n_peaks:
        L = int(rng.integers(lens[0], lens[1] + 1))
        pick = rng.choice(plain, size=L)
        want_mod = rng.random(L) < 0.15
        if want_mod.sum() > L // 2:
            drop = rng.choice(np.nonzero(want_mod)[0], size=int(want_mod.sum() - L // 2), replace=False)
            want_mod[drop] = False
        pick = np.where(want_mod, rng.choice(mods, size=L), pick)
        w = ims[pick] * 1e-3
        longest = max(longest, float(w.sum()))
        for ladder in (0, 1):
            su = np.cumsum(w if ladder == 0 else w[::-1])
            off = (off5 if ladder == 0 else off3)
            if full_dict and rng.random() < 0.5:
                off = (off5_alt if ladder == 0 else off3_alt)
            obs = (su + off) * (1 + rng.uniform(-ppm / 2, ppm / 2, size=L))
            if full_dict and ladder == 0 and off == off5_alt:
                # C4's "a-B" ions: the reference does not model them — an a-type fragment that has lost its
                # last nucleobase, i.e. the a/w offset minus a neutral base mass.  They are decoys: mostly invalid under
                # every breakage offset, and their differences to the neighbouring rungs explain nothing.
                decoy = rng.random(L) < 0.08
                obs = np.where(decoy, obs - rng.choice(NEUTRAL_BASE_MASSES, size=L), obs)
            else:
                decoy = np.zeros(L, dtype=bool)
            all_obs.append(obs)
            all_decoy.append(decoy)
            # validity: every peak against every breakage offset
            cand = obs[:, None] - offsets[None, :] * 1e-3
            v_mass.append(cand.ravel())
            v_thr.append(np.repeat(ppm * obs, len(offsets)))
            # explanation: adjacent kept rungs
            kept = np.nonzero(rng.random(L) >= miss_p)[0]
            su_obs = obs - off
            if len(kept) and kept[0] == 0:
                e_mass.append(np.array([su_obs[0]]))
                e_thr.append(np.array([ppm * obs[0]]))
                e_nt.append(np.array([1]))
                e_decoy.append(decoy[:1])
            if len(kept) > 1:
                gap = np.diff(kept)
                ok = gap <= max_gap
                a, b = kept[:-1][ok], kept[1:][ok]
                e_mass.append(su_obs[b] - su_obs[a])
                e_thr.append(ppm * (obs[a] + obs[b]))
                e_nt.append(gap[ok])
                e_decoy.append(decoy[a] | decoy[b])
            peaks += L
    min_w = float(ims[0]) * 1e-3
    max_len = int(longest / min_w)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dtype=dt)  # noqa: E731
    n_off = len(offsets)
    wl = Workload(name=config, config_id=cid, ppm=ppm, alphabet=reps, max_len=max_len, max_seq_length=msl,
                  n_peaks=peaks, valid_mass=cat(v_mass, np.float64), valid_thr=cat(v_thr, np.float64),
                  explain_mass=cat(e_mass, np.float64), explain_thr=cat(e_thr, np.float64),
                  explain_nt=cat(e_nt, np.int64), max_modifications=round(0.5 * max_len),
                  observed=cat(all_obs, np.float64), breakage={int(k): breakage[int(k)] for k in offsets},
                  observed_decoy=cat(all_decoy, bool), explain_decoy=cat(e_decoy, bool))
    # trim to exactly n_peaks peaks (validity arrays are peak-major)
    if peaks > n_peaks:
        wl.valid_mass = wl.valid_mass[: n_peaks * n_off]
        wl.valid_thr = wl.valid_thr[: n_peaks * n_off]
        wl.observed = wl.observed[:n_peaks]
        wl.observed_decoy = wl.observed_decoy[:n_peaks]
        wl.n_peaks = n_peaks
    return wl


def shard(wl: Workload, rank: int, world: int) -> Workload:
    """Interleaved shard of a workload (peaks are independent: no data-path collective)."""
    n_off = len(wl.valid_mass) // max(wl.n_peaks, 1)
    vm = wl.valid_mass.reshape(-1, n_off)[rank::world]
    vt = wl.valid_thr.reshape(-1, n_off)[rank::world]
    return Workload(name=wl.name, config_id=wl.config_id, ppm=wl.ppm, alphabet=wl.alphabet, max_len=wl.max_len,
                    max_seq_length=wl.max_seq_length, n_peaks=len(vm), valid_mass=vm.ravel(), valid_thr=vt.ravel(),
                    explain_mass=wl.explain_mass[rank::world], explain_thr=wl.explain_thr[rank::world],
                    explain_nt=wl.explain_nt[rank::world], max_modifications=wl.max_modifications,
                    observed=None if wl.observed is None else wl.observed[rank::world], breakage=wl.breakage,
                    observed_decoy=None if wl.observed_decoy is None else wl.observed_decoy[rank::world],
                    explain_decoy=None if wl.explain_decoy is None else wl.explain_decoy[rank::world])
