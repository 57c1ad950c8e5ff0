#!/usr/bin/env python3
"""Headline benchmark: peaks explained per second on synthetic spectra (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4] [--peaks 100000]
    python bench.py --impl reference ...        # the CPU restatement of the reference, all host cores

One step = one pass of the mass-explanation hot path over one batch: validity probes for every
(peak x breakage offset) + enumeration for every ladder difference of the batch.
`value` times the kernels with inputs resident in HBM; `e2e` times the public Python API with host
buffers (H2D + kernels + D2H of all results), four batches in flight on four context slots (the copies of
one under the kernels of the other).  Multi-GPU: one process per GPU (torchrun).  Default `--scaling weak`:
every rank explains its own 10^5-peak batch (no data-path collective), time = max over ranks.
`--scaling strong`: ONE fixed workload (the C4 batch tiled `--strong-factor` times) is partitioned over the
ranks in contiguous blocks of equal estimated work, every rank runs its block, and rank 0 gathers status /
offsets / records through POSIX shared memory inside the e2e region (spectrseqtools_b200/sharding.py).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import pathlib
import statistics
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "peaks_explained_per_sec"
UNIT = "peaks/s"


def ncu_traffic(kernel: str, wl=None):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json), or None.
    Only valid for the workload the capture was taken on (C4, 10^5 peaks)."""
    if wl is not None and not (wl.name == "C4" and wl.n_peaks == 100_000):
        return None
    try:
        doc = json.loads((ROOT / "profiles" / "traffic.json").read_text())
        return int(_kernel_entry(doc, kernel)["dram_bytes_per_launch"])
    except Exception:
        return None


def _kernel_entry(doc, kernel: str):
    """The entry of `kernel` in traffic.json: exact name, else the one instance whose name starts with it."""
    ks = doc["kernels"]
    if kernel in ks:
        return ks[kernel]
    hits = [k for k in ks if k.startswith(kernel)]
    return ks[hits[0]] if len(hits) == 1 else {}


def ncu_metric(kernel: str, key: str, wl=None):
    """Another per-kernel number of the same committed capture (e.g. issue_slot_util, a fraction), or None."""
    if wl is not None and not (wl.name == "C4" and wl.n_peaks == 100_000):
        return None
    try:
        doc = json.loads((ROOT / "profiles" / "traffic.json").read_text())
        return _kernel_entry(doc, kernel).get(key)
    except Exception:
        return None


def peaks_file():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.active = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in names.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(self.period)  # (the e2e region lowers the rate: NVML queries take driver-wide locks)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU legs (oracle = checker/baseline only)
_G = {}


def _cpu_init(weights, is_mod, rates, max_len, tol, max_mods, mode):
    """mode: "ref" = the reference's own function bodies (oracle/ref_harness.py, read from baseline/_ref or
    /root/reference), "py" = oracle/oracle_py.py, "c" = oracle/oracle.c.  The table always comes from the C
    restatement (bit-identical to set_up_bit_table by the SHA-256 goldens; the reference takes 200 s to build it)."""
    from oracle import oracle_c, oracle_py

    _G["w"] = list(weights)
    _G["tab"] = oracle_c.build_bit_table(list(weights), max(weights) * _G.get("msl", 35), 32)
    _G["rows"] = [oracle_py.Row(m, bool(im), rt) for m, im, rt in zip(weights, is_mod, rates)]
    _G["ind"] = oracle_py.individual_budgets(_G["rows"], max_len)
    _G.update(max_len=max_len, tol=tol, mm=max_mods, mode=mode)
    if mode == "ref":
        from oracle import ref_harness as H
        from spectrseqtools_b200 import masses as M

        ref = H.load_reference(dict(M._INT_MASS_NAMES), dict(M._INT_MASS_IS_MOD))
        masses = [ref.NucleotideMass(int(m), [], bool(im), float(rt)) for m, im, rt in zip(weights, is_mod, rates)]
        seq = ref.SequenceInformation(max_len=max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
        _G["ref"] = ref
        _G["duck"] = H.DuckTable(_G["tab"], masses, seq, precision=1e-3, tolerance=tol, compression_per_cell=32)


def _cpu_chunk(args):
    """Explain + validity (+ singleton test of every valid copy, as classify_fragments asks) for a chunk of calls."""
    from oracle import oracle_c, oracle_py

    e_mass, e_thr, v_mass, v_thr = args
    n = 0
    tab, rows, w, mode = _G["tab"], _G["rows"], _G["w"], _G["mode"]
    is_mod = [r.is_modification for r in rows]
    if mode == "ref":
        ref, duck = _G["ref"], _G["duck"]
        for m, t in zip(e_mass, e_thr):
            found = ref.explain_mass_with_table(float(m), duck, max_modifications=_G["mm"], threshold=float(t)).explanations
            n += 0 if found is None else len(found)
        for m, t in zip(v_mass, v_thr):
            try:
                if ref.is_valid_mass(float(m), duck, float(t)):
                    ref.is_singleton(float(m), w, duck, float(t))
            except NotImplementedError:
                pass
        return n
    for m, t in zip(e_mass, e_thr):
        if mode == "c":
            tg, th = oracle_py.integerise(float(m), float(t), 1e-3, _G["tol"])
            r, off, _ = oracle_c.explain(tab, 32, w, is_mod, _G["ind"], tg, th, _G["mm"], True)
            n += len(off) - 1
        else:
            n += len(oracle_py.explain_solutions(float(m), tab, rows, _G["max_len"], 32, 1e-3, _G["tol"], _G["mm"], float(t), True))
    for m, t in zip(v_mass, v_thr):
        try:
            if mode == "c":
                tg, th = oracle_py.integerise(float(m), float(t), 1e-3, _G["tol"])
                oracle_c.is_valid(tab, 32, tg, th)
            elif oracle_py.is_valid_mass(float(m), tab, 32, 1e-3, _G["tol"], float(t)):
                oracle_py.is_singleton(float(m), w, 1e-3, _G["tol"], float(t))  # classify_fragments asks this of every valid copy
        except NotImplementedError:
            pass
    return n


def reference_mode() -> str:
    """"ref" when the reference's sources are on this machine (baseline/_ref travels to the GPU box), else the port."""
    try:
        from oracle import ref_harness as H

        return "ref" if H.available() else "py"
    except Exception:
        return "py"


def cpu_leg(wl, table_rows, n_sample_peaks: int, mode: str, workers: int):
    """Time the CPU restatement on the first n_sample_peaks peaks of the workload (and their share of the
    explanation calls), spread over `workers` processes.  Returns (peaks/s, compositions, seconds)."""
    import multiprocessing as mp

    weights = [r[0] for r in table_rows]
    is_mod = [r[1] for r in table_rows]
    rates = [r[2] for r in table_rows]
    n_off = len(wl.valid_mass) // wl.n_peaks
    frac = n_sample_peaks / wl.n_peaks
    n_e = max(1, int(round(len(wl.explain_mass) * frac)))
    rng = np.random.default_rng(1)
    e_idx = np.sort(rng.choice(len(wl.explain_mass), size=n_e, replace=False))  # stratified by construction (random)
    v_idx = np.arange(n_sample_peaks * n_off)
    _G["msl"] = wl.max_seq_length
    chunks = []
    n_chunks = workers * 4
    for c in range(n_chunks):
        ei, vi = e_idx[c::n_chunks], v_idx[c::n_chunks]
        chunks.append((wl.explain_mass[ei], wl.explain_thr[ei], wl.valid_mass[vi], wl.valid_thr[vi]))
    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=_cpu_init, initargs=(weights, is_mod, rates, wl.max_len, wl.ppm, wl.max_modifications, mode)) as pool:
        pool.map(_cpu_chunk, [(np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0))] * workers)  # tables built, workers warm
        t0 = time.perf_counter()
        comps = sum(pool.map(_cpu_chunk, chunks))
        dt = time.perf_counter() - t0
    return n_sample_peaks / dt, comps, dt, n_e


def table_rows_for(wl):
    """(weight, is_mod, rate) per table row for the workload's alphabet with the universal 0.5 rate cap."""
    from spectrseqtools_b200 import masses as M
    from spectrseqtools_b200 import synthetic as S

    ims, reps, is_mod = S._rows(None if len(wl.alphabet) == 104 else wl.alphabet)
    rows = [(0, False, 0.0)]
    for m, mod in zip(ims, is_mod):
        rows.append((int(m), bool(mod), 0.5 if mod else 1.0))
    return rows


# ----------------------------------------------------------------------------- reference arm
def run_reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path on all host cores (rank 0 only): explain_mass_with_table /
    is_valid_mass / is_singleton bodies from baseline/_ref through oracle/ref_harness.py (kind "reference"); the
    port (oracle/oracle_py.py) only when the reference sources are not on this machine."""
    if rank != 0:
        return
    from spectrseqtools_b200 import synthetic as S

    wl = S.make_workload(args.workload, args.peaks)
    rows = table_rows_for(wl)
    workers = os.cpu_count() or 1
    mode = reference_mode()
    # bounded sample per step: sized from a quick calibration so that steps+warmup stay within ~2 minutes
    rate0, _, _, _ = cpu_leg(wl, rows, max(50, min(400, wl.n_peaks)), mode, workers)
    budget = 90.0 / max(1, args.steps + args.warmup)
    n_sample = int(max(100, min(wl.n_peaks, rate0 * budget)))
    for _ in range(args.warmup):
        cpu_leg(wl, rows, n_sample, mode, workers)
    total_t, comps = 0.0, 0
    for _ in range(args.steps):
        _r, c, dt, n_e = cpu_leg(wl, rows, n_sample, mode, workers)
        total_t += dt
        comps += c
    value = n_sample * args.steps / total_t
    sample = f"{n_sample} of {wl.n_peaks} peaks per step: {n_sample * (len(wl.valid_mass) // wl.n_peaks)} validity + ~{n_e} explanation calls"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(wl, args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "reference" if mode == "ref" else "port", "sample": sample,
                         "what": CPU_WHAT[mode]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    line["explanations_per_sec" if mode == "ref" else "compositions_per_sec"] = comps / total_t
    print(json.dumps(line))


CPU_WHAT = {
    "ref": "the reference's unmodified explain_mass_with_table / is_valid_mass / is_singleton bodies (baseline/_ref, loaded by "
           "oracle/ref_harness.py; pure Python like the reference), table from the C restatement (SHA-256 pinned)",
    "py": "oracle/oracle_py.py: pure-Python restatement of the reference functions (reference sources not on this machine)",
    "c": "oracle/oracle.c (same algorithm in C)",
}


def workload_config(wl, args):
    return {"workload": f"{wl.name}: {args.peaks} synthetic ladder peaks per GPU, {len(wl.alphabet)}-mass alphabet, "
                        f"{wl.ppm * 1e6:g} ppm; per step {len(wl.valid_mass)} validity probes + {len(wl.explain_mass)} explanation calls",
            "peaks_per_gpu": wl.n_peaks, "l2": "flushed between timed steps (256 MiB write, untimed)",
            "table_rows": len(wl.alphabet) + 1, "max_len": wl.max_len, "max_modifications": wl.max_modifications}


def spread(xs):
    xs = sorted(xs)
    q = lambda f: xs[min(len(xs) - 1, int(f * (len(xs) - 1) + 0.5))]  # noqa: E731
    return {"median": q(0.5), "p10": q(0.1), "p90": q(0.9), "min": xs[0], "max": xs[-1]}


def tile_workload(wl, f):
    """The same batch `f` times over (one fixed larger workload: the throughput regime / the strong-scaling job)."""
    import copy

    big = copy.copy(wl)
    big.n_peaks = wl.n_peaks * f
    for k in ("valid_mass", "valid_thr", "explain_mass", "explain_thr", "explain_nt", "observed"):
        setattr(big, k, np.tile(getattr(wl, k), f))
    return big


def block_of(wl, rank, world, dp):
    """Strong scaling: rank's contiguous block of the fixed workload (explanation calls in blocks of equal estimated
    work, observed peaks in equal blocks — validity cost is uniform)."""
    import copy

    from spectrseqtools_b200 import sharding

    # compositions per call looked up on the device (every rank asks for the whole workload: 10^6 calls take well under a
    # millisecond) — blocks of equal OUTPUT, known before anything is enumerated
    from spectrseqtools_b200 import mass_explanation as ME

    counts = ME.count_compositions(wl.explain_mass, dp, wl.explain_thr)
    eb = sharding.partition_contiguous(wl.explain_mass, wl.explain_thr, world, dp, counts=counts)
    n_off = len(wl.valid_mass) // max(wl.n_peaks, 1)
    ob = [(wl.n_peaks * r) // world for r in range(world + 1)]
    mine = copy.copy(wl)
    lo, hi = eb[rank], eb[rank + 1]
    mine.explain_mass, mine.explain_thr, mine.explain_nt = wl.explain_mass[lo:hi], wl.explain_thr[lo:hi], wl.explain_nt[lo:hi]
    mine.observed = wl.observed[ob[rank]:ob[rank + 1]]
    mine.valid_mass = wl.valid_mass[ob[rank] * n_off:ob[rank + 1] * n_off]
    mine.valid_thr = wl.valid_thr[ob[rank] * n_off:ob[rank + 1] * n_off]
    mine.n_peaks = ob[rank + 1] - ob[rank]
    return mine, (lo, hi), (ob[rank], ob[rank + 1])


# ----------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--peaks", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--large-factor", type=int, default=16,
                    help="also time the same step on the batch tiled this many times (throughput regime; 0 = skip)")
    ap.add_argument("--flush", default="write", choices=["write", "none"],
                    help="L2 between timed steps: write a 256 MiB buffer (default, the contract) or leave it warm (diagnostics)")
    ap.add_argument("--e2e-depth", type=int, default=4, help="batches in flight in the e2e loop (context slots)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="multi-GPU: one batch per GPU (weak, default) or ONE fixed workload partitioned over the ranks and gathered on rank 0 (strong)")
    ap.add_argument("--strong-factor", type=int, default=16, help="strong scaling: the fixed workload is the batch tiled this many times")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world > 1:  # one rank per GPU, each on its own cores (32 host cores, all on one NUMA node, shared by 8 ranks)
        try:
            n_cpu = os.cpu_count() or 1
            per = max(1, n_cpu // world)
            os.sched_setaffinity(0, set(range(local_rank * per, (local_rank + 1) * per)))
        except (AttributeError, OSError):
            pass

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from spectrseqtools_b200 import _cabi
    from spectrseqtools_b200 import fragment_classification as FC
    from spectrseqtools_b200 import mass_explanation as ME
    from spectrseqtools_b200 import mass_table as MT
    from spectrseqtools_b200 import sharding
    from spectrseqtools_b200 import synthetic as S

    peaks_doc, peak_src = peaks_file()
    hbm_peak = float(peaks_doc.get("hbm_gbs", 6650.0))
    ctx = _cabi.context(local_rank)
    strong = args.scaling == "strong"
    wl_all = S.make_workload(args.workload, args.peaks, seed_offset=0 if strong else rank)
    MT.MAX_SEQ_LENGTH = wl_all.max_seq_length
    seq = MT.SequenceInformation(max_len=wl_all.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
    frame = S.alphabet_frame(None if len(wl_all.alphabet) == 104 else wl_all.alphabet)
    dp = MT.DynamicProgrammingTable(frame, 32, wl_all.ppm, 1e-3, seq, device=local_rank)
    dev = dp.device_table()
    e_block = o_block = None
    if strong:
        wl_all = tile_workload(wl_all, args.strong_factor)
        wl, e_block, o_block = block_of(wl_all, rank, world, dp)
    else:
        wl = wl_all

    # ---- table build (K1 + K1t), timed alone: burst
    build_ms, tr_ms = [], []
    for _ in range(5):
        ctx.flush_l2()
        dev.rebuild()
        b, t = dev.timings()
        build_ms.append(b)
        tr_ms.append(t)
    table_bytes = dev.R * dev.C * 8
    mask_bytes = dev.C * 32 * 16
    full_c4 = dev.R == 105 and wl.max_seq_length == 35
    table_info = {
        "rows": dev.R, "words_per_row": dev.C, "bytes": table_bytes,
        "build_ms": min(build_ms), "build_ms_median": statistics.median(build_ms),
        "roofline": {"bound": "hbm", "achieved": table_bytes / (min(build_ms) * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": table_bytes / (min(build_ms) * 1e-3) / 1e9 / hbm_peak, "traffic": ncu_traffic("k_build_table<8, 0>") if full_c4 else None,
                     "algorithmic_bytes": table_bytes, "note": "R*C*8 bytes written once (SURVEY §8d K1)"},
        "row_masks_ms": min(tr_ms),
        "row_masks_roofline": {"bound": "hbm", "achieved": (table_bytes + mask_bytes) / (min(tr_ms) * 1e-3) / 1e9, "peak": hbm_peak,
                               "unit": "GB/s", "frac": (table_bytes + mask_bytes) / (min(tr_ms) * 1e-3) / 1e9 / hbm_peak,
                               "traffic": ncu_traffic("k_transpose_masks") if full_c4 else None,
                               "algorithmic_bytes": table_bytes + mask_bytes},
    }
    extra = dev.extra_timings() if hasattr(dev, "extra_timings") else None
    if extra:
        table_info.update(extra)

    # ---- stage the batch (inputs resident in HBM for `value`); host inputs live in pinned memory
    def pinned_copy(a):
        out = ctx.pinned_empty(a.shape, a.dtype)
        out[...] = a
        return out

    observed = pinned_copy(wl.observed)
    offsets = np.array([w * dp.precision for w in wl.breakage], dtype=np.float64)
    e_mass, e_thrf = pinned_copy(wl.explain_mass), pinned_copy(wl.explain_thr)
    v_target, v_thr = ME._integerise_many(wl.valid_mass, wl.valid_thr, dp)      # only for the byte accounting below
    e_target, e_thr = ME._integerise_many(wl.explain_mass, wl.explain_thr, dp)
    weights, is_mod, ind = ME._row_metadata(dp)
    max_mods = np.full(len(e_target), wl.max_modifications, dtype=np.int32)
    ctx.classify_stage(observed, offsets)   # validity of every (peak x breakage offset): the fused N2 front end
    ctx.explain_stage_f64(dev, e_mass, e_thrf, max_mods, ind, is_mod, dp.precision, dp.tolerance, True)

    def step():
        ctx.classify_launch(dev, dp.precision, dp.tolerance)  # queued; explain_run's synchronisation completes both
        return ctx.explain_run(dev, 0)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        n_roots, n_comps = step()
    if dist is not None:
        dist.barrier()
    ctx.stats_reset()
    sampler.active.set()
    step_ms = []
    for _ in range(args.steps):
        if args.flush == "write":
            ctx.flush_l2()
        ctx.timer_start()
        n_roots, n_comps = step()
        step_ms.append(ctx.timer_stop_at_run())  # CUDA events on the launching stream: start of the step .. last device op of it
    sampler.active.clear()
    total_ms = float(sum(step_ms))
    stats = ctx.kernel_stats()
    pass_used = ctx.last_pass()
    if dist is not None:
        import torch

        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        cnt = torch.tensor([float(wl.n_peaks), float(n_comps)], device="cuda", dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        peaks_all, comps_all = float(cnt[0].item()), float(cnt[1].item())
    else:
        peaks_all, comps_all = float(wl.n_peaks), float(n_comps)
    value = peaks_all * args.steps / (total_ms * 1e-3)

    # ---- e2e: public API, host buffers in, host arrays out (H2D + kernels + D2H of every result), four batches in
    # flight on four context slots: steps i+1 .. i+3 are submitted before step i is collected, so the copies of one batch run
    # under the kernels of the other.  Strong scaling: every rank publishes its block in shared memory and rank 0
    # gathers all blocks inside the timed region.
    gather = sharding.ShmGather(f"sstb200_{os.environ.get('MASTER_PORT', '0')}", rank, world, capacity=1 << 30) if (strong and world > 1) else None
    if gather is not None:
        gather.register(ctx)  # page-locked: the device-to-host copies land in the shared segment, the gather copies nothing
    flag_bytes = (len(wl.breakage) * ((len(observed) + 1) // 2) + 4095) & ~4095

    def submit(slot, seq_no):
        out = block = None
        if gather is not None:
            reg = gather.region(seq_no)
            out, block = reg[:flag_bytes], reg[flag_bytes:]
        v = FC.classify_observed(observed, dp, wl.breakage, copy=False, wait=False, slot=slot, out=out)
        b = ME.explain_masses(e_mass, dp, max_modifications=wl.max_modifications, thresholds=e_thrf, copy=False, wait=False, slot=slot,
                              out_block=block)
        return v, b

    def finish(pend, seq_no):
        v, b = pend
        v.wait()
        batch = b.wait()
        if gather is not None:
            gather.publish(seq_no, [batch.status, batch._offsets] + batch.raw_records() + [v._flags])
            if rank == 0:
                return v, batch, gather.collect(seq_no)
        return v, batch, None

    E2E_DEPTH = args.e2e_depth  # batches in flight: four context slots keep the copy engines and the SMs busy at the same time (a batch's
    # chain — submit, copy in, stage, pass, copy out, collect — is ~310 us long: 137 us per step with three in flight, 128
    # with four, 120 with six, where the submitting host thread becomes the limit)

    assert gather is None or E2E_DEPTH + 2 <= gather.REGIONS

    def e2e_loop(n, seq0):
        pend = [submit(k, seq0 + k) for k in range(min(E2E_DEPTH - 1, n))]
        out = None
        for i in range(n):
            if i + E2E_DEPTH - 1 < n:
                pend.append(submit((i + E2E_DEPTH - 1) % E2E_DEPTH, seq0 + i + E2E_DEPTH - 1))
            out = finish(pend.pop(0), seq0 + i)
        return out

    e2e_loop(2 * E2E_DEPTH, 1)  # (the gather numbers its steps consecutively: the timed loop continues at 2 * E2E_DEPTH + 1)
    if dist is not None:
        dist.barrier()
    sampler.period = 0.05
    sampler.active.set()
    t0 = time.perf_counter()
    valid, batch, gathered = e2e_loop(args.steps, 2 * E2E_DEPTH + 1)
    e2e_s = time.perf_counter() - t0
    sampler.active.clear()
    sampler.stop()
    if dist is not None:
        import torch

        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    rec_width = batch.records.shape[1] if batch.records.ndim == 2 else 8
    h2d = 8 * len(observed) + 8 * len(offsets) + 16 * len(e_target) + 5 * dev.R  # masses + thresholds; the batch-wide budget is a scalar
    # flags (two per byte) + the result block of the enumeration (header, status, uint32 offsets, records incl. the copy margin)
    d2h = int(valid._flags.size) + _cabi.context(local_rank, (args.steps - 1) % E2E_DEPTH).explain_d2h_bytes()
    e2e_value = peaks_all * args.steps / e2e_s

    # ---- the reference-shaped scalar entries (one mass per call: what prediction.py / skeleton_building.py do)
    scalar = None
    if rank == 0:
        k = min(200, len(wl.explain_mass))
        idx = np.linspace(0, len(wl.explain_mass) - 1, k).astype(int)
        for _ in range(2):
            ME.explain_mass_with_table(float(wl.explain_mass[idx[0]]), dp, wl.max_modifications, threshold=float(wl.explain_thr[idx[0]]))
        te, tv = [], []
        for i in idx:
            t1 = time.perf_counter()
            ME.explain_mass_with_table(float(wl.explain_mass[i]), dp, wl.max_modifications, threshold=float(wl.explain_thr[i]))
            t2 = time.perf_counter()
            try:
                ME.is_valid_mass(float(wl.valid_mass[i]), dp, float(wl.valid_thr[i]))
            except NotImplementedError:  # a probe beyond the table (C5's heaviest fragments): the reference raises too
                pass
            t3 = time.perf_counter()
            te.append((t2 - t1) * 1e6)
            tv.append((t3 - t2) * 1e6)
        scalar = {"explain_mass_with_table": spread(te), "is_valid_mass": spread(tv), "calls": int(k),
                  "note": "wall time of one reference-shaped call (a batch of one), names included"}

    # ---- rows N3 / N4 on the device: one round of the explanation-based alphabet reduction over a fragment frame that
    # stays on the device (pairs + thresholds generated, enumerated, deduplicated by key, rows united, fragments
    # re-validated); wall time of the calls a Predictor.filter_by_explanation round makes
    ladder = None
    if rank == 0 and not strong:
        from spectrseqtools_b200 import alphabet_reduction as AR

        ladder = {}
        for tag, n_oligos in (("one_spectrum", 3), ("merged_100_oligos", 100)):
            su, obs_f, brk, single = S.make_frame(n_oligos, names=None if len(wl.alphabet) == 104 else wl.alphabet)
            lad = AR.DeviceLadder(su, obs_f, brk, single, dp, S.alphabet_frame(None if len(wl.alphabet) == 104 else wl.alphabet))
            lad.round(); lad.revalidate()
            tr, tv = [], []
            for _ in range(20):
                t1 = time.perf_counter()
                lad.round()
                t2 = time.perf_counter()
                lad.revalidate()
                t3 = time.perf_counter()
                tr.append((t2 - t1) * 1e6)
                tv.append((t3 - t2) * 1e6)
            ladder[tag] = {"fragments": int(len(su)), "calls": int(lad.n_calls), "compositions": int(lad.n_compositions),
                           "round_us": spread(tr), "revalidate_us": spread(tv)}
        ladder["note"] = ("sst_ladder_round / sst_ladder_revalidate: wall time per call; per round the host receives a 128-bit row mask "
                          "and two counters (the table rebuild between them is table_build)")

    # ---- roofline of the enumeration pass (K2a + K2b family) from the live per-kernel event times
    win_words = ((2 * e_thr + 1 + 31) // 32 + 1).sum()
    comp_len = int((batch.records > 0).sum()) if batch.records is not None and batch.records.size else 0
    k2b_bytes = 16 * len(e_target) + 8 * (len(e_target) + 1) + 4 * int(batch.n_compositions) + comp_len + 8 * int(win_words)
    vwin_words = ((2 * v_thr + 1 + 31) // 32 + 1).sum()
    k2a_bytes = 8 * len(observed) + len(v_target) + 8 * int(vwin_words)
    fam = ["phase_a", "explain_pass"]
    k2b_ms = sum(stats[k][0] for k in fam) / args.steps
    k2a_ms = stats["classify"][0] / args.steps
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in stats.items() if v[1]}
    dominant = max(fam + ["classify"], key=lambda k: stats[k][0])
    pass_name = {1: "k_explain_pass (level-synchronous)", 2: "k_explain_dfs (count -> scan -> fill over items in peak order)",
                 3: "k_explain_direct (counts from the composition-count table -> scan -> output-balanced fill)"}.get(pass_used, str(pass_used))
    tkey = {1: "k_explain_pass<1>", 2: "k_explain_dfs<1, 0", 3: "k_explain_direct<1>"}.get(pass_used, "")
    roofline = {"bound": "hbm", "kernel": "K3+K2b enumeration pass: " + pass_name,
                "achieved": k2b_bytes / (k2b_ms * 1e-3) / 1e9 if k2b_ms else None, "peak": hbm_peak, "unit": "GB/s",
                "frac": (k2b_bytes / (k2b_ms * 1e-3) / 1e9 / hbm_peak) if k2b_ms else None,
                "traffic": ncu_traffic(tkey, wl), "traffic_source": "profiles/traffic.json (ncu --set full, same workload)",
                "algorithmic_bytes": int(k2b_bytes), "peak_source": peak_src, "dominant_launch": dominant,
                "k2a": {"kernel": "k_classify (validity + singleton of every peak x breakage pair)",
                        "note": "answers most windows from an L1-resident summary: issue/latency bound, NOT a DRAM stream — algorithmic_bytes are the "
                                "window words SURVEY §8d counts, dram traffic is what ncu measured; see issue_slot_util",
                        "algorithmic_bytes": int(k2a_bytes), "ms": k2a_ms, "traffic": ncu_traffic("k_classify", wl),
                        "issue_slot_util": ncu_metric("k_classify", "issue_slot_util", wl)}}
    launches = sum(v[1] for v in stats.values())
    ph = ctx.explain_phase_ns().astype(np.int64)
    ph = ph[ph > 0]
    phase_us = [round(float(x) * 1e-3, 2) for x in np.diff(ph)] if len(ph) > 1 else None  # see sst_explain_phase_ns

    # ---- the same step on a batch `large_factor` times larger (rank 0): where the latency of the dependent phases
    # of the pass is amortised and bytes per second mean something
    large = None
    if rank == 0 and args.large_factor > 1 and not strong:
        f = args.large_factor
        big_obs = pinned_copy(np.tile(wl.observed, f))
        big_mass, big_thr = pinned_copy(np.tile(wl.explain_mass, f)), pinned_copy(np.tile(wl.explain_thr, f))
        big_mm = np.full(len(big_mass), wl.max_modifications, dtype=np.int32)
        ctx.classify_stage(big_obs, offsets)
        ctx.explain_stage_f64(dev, big_mass, big_thr, big_mm, ind, is_mod, dp.precision, dp.tolerance, True)
        for _ in range(3):
            _r, big_comps = step()
        ctx.stats_reset()
        big_ms, big_steps = 0.0, 10
        for _ in range(big_steps):
            if args.flush == "write":
                ctx.flush_l2()
            ctx.timer_start()
            _r, big_comps = step()
            big_ms += ctx.timer_stop_at_run()
        bstats = ctx.kernel_stats()
        pass_ms = bstats["explain_pass"][0] / big_steps
        cls_ms = bstats["classify"][0] / big_steps
        large = {"factor": f, "peaks": wl.n_peaks * f, "ms_per_step": big_ms / big_steps,
                 "peaks_per_sec": wl.n_peaks * f * big_steps / (big_ms * 1e-3),
                 "compositions_per_sec": big_comps * big_steps / (big_ms * 1e-3),
                 "explain_pass_ms": pass_ms, "classify_ms": cls_ms,
                 "explain_pass_gbs": k2b_bytes * f / (pass_ms * 1e-3) / 1e9, "explain_pass_frac": k2b_bytes * f / (pass_ms * 1e-3) / 1e9 / hbm_peak,
                 "classify_pairs_per_sec": len(v_target) * f / (cls_ms * 1e-3)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": workload_config(wl, args),
        "compositions_per_sec": comps_all * args.steps / (total_ms * 1e-3), "compositions_per_step": comps_all,
        "roots_per_step_rank0": int(n_roots),
        "ms_per_step_spread_rank0": spread(step_ms),
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "how": "public API (classify_observed + explain_masses, wait=False), pinned host inputs, results in pinned host "
                       "arrays; four batches in flight on four context slots" + ("; every rank's results are copied by the device straight into its page-locked POSIX shared-memory segment, which rank 0 maps: gathered inside the timed region without a host copy" if gather is not None else "")},
        "scalar_latency_us": scalar,
        "ladder_round": ladder,
        "gpu_launches": int(launches),
        "roofline": roofline, "kernels": kernels, "pass_phase_us": phase_us, "large_batch": large, "table_build": table_info,
    }
    if strong:
        line["config"]["workload"] = (f"{wl_all.name} x{args.strong_factor}: ONE fixed workload of {wl_all.n_peaks} peaks / {len(wl_all.explain_mass)} explanation calls, "
                                      f"partitioned over {world} rank(s) in contiguous blocks of equal estimated work, gathered on rank 0")
        line["config"]["peaks_total"] = wl_all.n_peaks
        if gathered is not None:
            line["strong_gather"] = {"blocks": len(gathered), "compositions": int(sum(int(g[1][-1]) for g in gathered))}

    # ranks other than 0 are done here: they leave before rank 0's CPU legs (otherwise they spin in an NCCL barrier
    # for half a minute and the driver's busy sample reads it as GPU work)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
        dist = None
    if gather is not None:
        gather.close()
    if rank != 0:
        return

    if not args.no_parity and not strong:
        line["parity"] = parity_gate(dp, wl, batch, valid, dev)
    if not args.no_cpu_baseline:
        rows = table_rows_for(wl)
        workers = os.cpu_count() or 1
        mode = reference_mode()
        rate0, _, _, _ = cpu_leg(wl, rows, max(50, min(400, wl.n_peaks)), mode, workers)
        n_sample = int(max(100, min(wl.n_peaks, rate0 * 15.0)))
        rate, comps, dt, n_e = cpu_leg(wl, rows, n_sample, mode, workers)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": workers, "kind": "reference" if mode == "ref" else "port",
                                "sample": f"first {n_sample} of {wl.n_peaks} peaks ({n_sample * (len(wl.valid_mass) // wl.n_peaks)} validity + {n_e} explanation calls), {dt:.1f} s wall",
                                "what": CPU_WHAT[mode]}
        n_c = int(min(wl.n_peaks, max(2000, n_sample * 20)))
        rate_c, comps_c, dt_c, n_ec = cpu_leg(wl, rows, n_c, "c", workers)
        line["cpu_baseline_c"] = {"value": rate_c, "unit": UNIT, "cores": workers, "kind": "port",
                                  "sample": f"first {n_c} peaks, {dt_c:.1f} s wall", "what": CPU_WHAT["c"],
                                  "compositions_per_sec": comps_c / dt_c}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- same-run parity gate (C oracle = checker)
def _gate_chunk(args):
    from oracle import oracle_c

    kind, a, b = args
    if kind == "e":
        return oracle_c.explain_batch_keys(_G["tab"], 32, _G["w"], _G["is_mod"], _G["ind"], a, b, _G["mm"], True)
    return oracle_c.is_valid_batch(_G["tab"], 32, a, b)


def parity_gate(dp, wl, batch, valid, dev):
    """Same-run parity: table SHA-256 against the golden value, then EVERY explanation call and EVERY validity probe of
    the step against the C oracle (all host cores, a few seconds)."""
    import multiprocessing as mp

    from oracle import oracle_c, oracle_py
    from spectrseqtools_b200 import mass_explanation as ME

    out = {}
    gold = json.loads((ROOT / "tests" / "golden" / "tables_sha.json").read_text())
    weights = [m.mass for m in dp.masses]
    if weights == gold["full"]["weights"] and dev.C == gold["full"]["shape"][1]:
        out["table_sha256_matches_reference"] = hashlib.sha256(dev.download().tobytes()).hexdigest() == gold["full"]["sha256"]
    tab = oracle_c.build_bit_table(weights, max(weights) * wl.max_seq_length, 32)
    if "table_sha256_matches_reference" not in out:
        out["table_equals_c_oracle"] = bool(np.array_equal(dev.download(), tab))
    rows = [oracle_py.Row(m.mass, m.is_modification, m.modification_rate) for m in dp.masses]
    _G.update(tab=tab, w=weights, is_mod=[r.is_modification for r in rows], ind=oracle_py.individual_budgets(rows, dp.seq.max_len),
              mm=wl.max_modifications)
    e_target, e_thr = ME._integerise_many(wl.explain_mass, wl.explain_thr, dp)
    v_target, v_thr = ME._integerise_many(wl.valid_mass, wl.valid_thr, dp)
    workers = os.cpu_count() or 1
    n_e, n_v = len(e_target), len(v_target)
    e_cuts = [(n_e * i) // (workers * 4) for i in range(workers * 4 + 1)]
    v_cuts = [(n_v * i) // (workers * 4) for i in range(workers * 4 + 1)]
    jobs = [("e", e_target[a:b], e_thr[a:b]) for a, b in zip(e_cuts[:-1], e_cuts[1:])]
    jobs += [("v", v_target[a:b], v_thr[a:b]) for a, b in zip(v_cuts[:-1], v_cuts[1:])]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(workers) as pool:
        res = pool.map(_gate_chunk, jobs)
    e_res, v_res = res[: workers * 4], res[workers * 4:]
    want_counts = np.concatenate([r[0] for r in e_res])
    want_keys = np.concatenate([r[1] for r in e_res])
    want_codes = np.concatenate(v_res)
    # device side: records as little-endian 8-byte keys, sorted inside every call
    recs = np.ascontiguousarray(batch.records)
    counts = batch.counts()
    ok_counts = bool(np.array_equal(np.where(want_counts < 0, 0, want_counts), counts))
    ok_status = bool(np.array_equal((batch.status & 2) != 0, want_counts < 0))
    ok_keys = False
    if recs.ndim == 2 and recs.shape[1] == 8 and ok_counts:
        keys = recs.view(np.uint64).reshape(-1)
        call = np.repeat(np.arange(len(counts)), counts)
        keys = keys[np.lexsort((keys, call))]
        ok_keys = bool(np.array_equal(keys, want_keys))
    out["explain_checked"] = int(n_e)
    out["explain_compositions_checked"] = int(len(want_keys))
    out["explain_ok"] = bool(ok_counts and ok_status and ok_keys)
    out["explain_digest"] = hashlib.sha256(want_keys.tobytes()).hexdigest()[:16]
    flags = valid.flags  # [breakage, peak]; wl.valid_* are peak-major
    n_off = len(wl.breakage)
    got_codes = np.ascontiguousarray(flags.T).reshape(-1) & 3
    out["validity_checked"] = int(n_v)
    out["validity_ok"] = bool(np.array_equal(got_codes[:n_v], want_codes))
    out["oracle_seconds"] = round(time.perf_counter() - t0, 2)
    return out


if __name__ == "__main__":
    main()
