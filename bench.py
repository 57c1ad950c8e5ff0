#!/usr/bin/env python3
"""Headline benchmark: peaks explained per second on synthetic spectra (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4] [--peaks 100000]
    python bench.py --impl reference ...        # the CPU restatement of the reference, all host cores

One step = one pass of the mass-explanation hot path over one batch: validity probes for every
(peak x breakage offset) + enumeration for every ladder difference of the batch.
`value` times the kernels with inputs resident in HBM; `e2e` times the public Python API with host
buffers (H2D + kernels + D2H of all results).  Multi-GPU: one process per GPU (torchrun), every rank
explains its own 10^5-peak batch (weak scaling, no data-path collective), time = max over ranks.
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
        return int(doc["kernels"][kernel]["dram_bytes_per_launch"])
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
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU legs (oracle = checker/baseline only)
_G = {}


def _cpu_init(weights, is_mod, rates, max_len, tol, max_mods, use_c):
    from oracle import oracle_c, oracle_py

    _G["w"] = list(weights)
    _G["tab"] = oracle_c.build_bit_table(list(weights), max(weights) * _G.get("msl", 35), 32)
    _G["rows"] = [oracle_py.Row(m, bool(im), rt) for m, im, rt in zip(weights, is_mod, rates)]
    _G["ind"] = oracle_py.individual_budgets(_G["rows"], max_len)
    _G.update(max_len=max_len, tol=tol, mm=max_mods, use_c=use_c)


def _cpu_chunk(args):
    """Explain + validity for a chunk of calls with the reference's algorithm (Python port or C port)."""
    from oracle import oracle_c, oracle_py

    e_mass, e_thr, v_mass, v_thr = args
    n = 0
    tab, rows, w = _G["tab"], _G["rows"], _G["w"]
    is_mod = [r.is_modification for r in rows]
    for m, t in zip(e_mass, e_thr):
        if _G["use_c"]:
            tg, th = oracle_py.integerise(float(m), float(t), 1e-3, _G["tol"])
            r, off, _ = oracle_c.explain(tab, 32, w, is_mod, _G["ind"], tg, th, _G["mm"], True)
            n += len(off) - 1
        else:
            n += len(oracle_py.explain_solutions(float(m), tab, rows, _G["max_len"], 32, 1e-3, _G["tol"], _G["mm"], float(t), True))
    for m, t in zip(v_mass, v_thr):
        try:
            if _G["use_c"]:
                tg, th = oracle_py.integerise(float(m), float(t), 1e-3, _G["tol"])
                oracle_c.is_valid(tab, 32, tg, th)
            elif oracle_py.is_valid_mass(float(m), tab, 32, 1e-3, _G["tol"], float(t)):
                oracle_py.is_singleton(float(m), w, 1e-3, _G["tol"], float(t))  # classify_fragments asks this of every valid copy
        except NotImplementedError:
            pass
    return n


def cpu_leg(wl, table_rows, n_sample_peaks: int, use_c: bool, workers: int):
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
    with ctx.Pool(workers, initializer=_cpu_init, initargs=(weights, is_mod, rates, wl.max_len, wl.ppm, wl.max_modifications, use_c)) as pool:
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
    if rank != 0:
        return
    from spectrseqtools_b200 import synthetic as S

    wl = S.make_workload(args.workload, args.peaks)
    rows = table_rows_for(wl)
    workers = os.cpu_count() or 1
    # bounded sample per step: sized from a quick calibration so that steps+warmup stay within ~2 minutes
    rate0, _, _, _ = cpu_leg(wl, rows, max(50, min(400, wl.n_peaks)), False, workers)
    budget = 90.0 / max(1, args.steps + args.warmup)
    n_sample = int(max(100, min(wl.n_peaks, rate0 * budget)))
    for _ in range(args.warmup):
        cpu_leg(wl, rows, n_sample, False, workers)
    total_t, comps = 0.0, 0
    for _ in range(args.steps):
        _r, c, dt, n_e = cpu_leg(wl, rows, n_sample, False, workers)
        total_t += dt
        comps += c
    value = n_sample * args.steps / total_t
    sample = f"{n_sample} of {wl.n_peaks} peaks per step: {n_sample * (len(wl.valid_mass) // wl.n_peaks)} validity + ~{n_e} explanation calls"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(wl, args),
        "compositions_per_sec": comps / total_t,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample,
                         "what": "oracle/oracle_py.py (pure-Python restatement of the reference's explain_mass_with_table / is_valid_mass; the reference itself is pure Python and cannot be imported here: polars missing)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(wl, args):
    return {"workload": f"{wl.name}: {args.peaks} synthetic ladder peaks per GPU, {len(wl.alphabet)}-mass alphabet, "
                        f"{wl.ppm * 1e6:g} ppm; per step {len(wl.valid_mass)} validity probes + {len(wl.explain_mass)} explanation calls",
            "peaks_per_gpu": wl.n_peaks, "l2": "flushed between timed steps (256 MiB write, untimed)",
            "table_rows": len(wl.alphabet) + 1, "max_len": wl.max_len, "max_modifications": wl.max_modifications}


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

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
    from spectrseqtools_b200 import synthetic as S

    peaks_doc, peak_src = peaks_file()
    hbm_peak = float(peaks_doc.get("hbm_gbs", 6650.0))
    ctx = _cabi.context(local_rank)
    wl = S.make_workload(args.workload, args.peaks, seed_offset=rank)
    MT.MAX_SEQ_LENGTH = wl.max_seq_length
    seq = MT.SequenceInformation(max_len=wl.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
    frame = S.alphabet_frame(None if len(wl.alphabet) == 104 else wl.alphabet)
    dp = MT.DynamicProgrammingTable(frame, 32, wl.ppm, 1e-3, seq, device=local_rank)
    dev = dp.device_table()

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
    table_info = {
        "rows": dev.R, "words_per_row": dev.C, "bytes": table_bytes,
        "build_ms": min(build_ms), "build_ms_median": statistics.median(build_ms),
        "roofline": {"bound": "hbm", "achieved": table_bytes / (min(build_ms) * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": table_bytes / (min(build_ms) * 1e-3) / 1e9 / hbm_peak, "traffic": ncu_traffic("k_build_table<8, 0>") if dev.R == 105 and wl.max_seq_length == 35 else None,
                     "algorithmic_bytes": table_bytes, "note": "R*C*8 bytes written once (SURVEY §8d K1)"},
        "row_masks_ms": min(tr_ms),
        "row_masks_roofline": {"bound": "hbm", "achieved": (table_bytes + mask_bytes) / (min(tr_ms) * 1e-3) / 1e9, "peak": hbm_peak,
                               "unit": "GB/s", "frac": (table_bytes + mask_bytes) / (min(tr_ms) * 1e-3) / 1e9 / hbm_peak,
                               "traffic": ncu_traffic("k_transpose_masks") if dev.R == 105 and wl.max_seq_length == 35 else None,
                               "algorithmic_bytes": table_bytes + mask_bytes},
    }

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
    total_ms = 0.0
    for _ in range(args.steps):
        if args.flush == "write":
            ctx.flush_l2()
        ctx.timer_start()
        n_roots, n_comps = step()
        total_ms += ctx.timer_stop_at_run()  # CUDA events on the launching stream: start of the step .. last device op of it
    sampler.active.clear()
    stats = ctx.kernel_stats()
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

    # ---- e2e: public API, host buffers in, host arrays out (H2D + kernels + D2H of every result)
    def e2e_step():
        valid = FC.classify_observed(observed, dp, wl.breakage, copy=False, wait=False)  # side stream: overlaps the enumeration
        batch = ME.explain_masses(e_mass, dp, max_modifications=wl.max_modifications, thresholds=e_thrf, copy=False)
        valid.wait()
        return valid, batch

    for _ in range(2):
        valid, batch = e2e_step()
    if dist is not None:
        dist.barrier()
    sampler.active.set()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        valid, batch = e2e_step()
    e2e_s = time.perf_counter() - t0
    sampler.active.clear()
    sampler.stop()
    if dist is not None:
        import torch

        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    rec_width = ctx._last[2]
    h2d = 8 * len(observed) + 8 * len(offsets) + 16 * len(e_target) + 5 * dev.R  # masses + thresholds; the batch-wide budget is a scalar
    d2h = len(v_target) + len(e_target) + 8 * (len(e_target) + 1) + int(batch.n_compositions) * rec_width
    e2e_value = peaks_all * args.steps / e2e_s

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
    roofline = {"bound": "hbm", "kernel": "K3+K2b enumeration pass (k_explain_pass: window roots -> items -> compositions, one cooperative launch)",
                "achieved": k2b_bytes / (k2b_ms * 1e-3) / 1e9 if k2b_ms else None, "peak": hbm_peak, "unit": "GB/s",
                "frac": (k2b_bytes / (k2b_ms * 1e-3) / 1e9 / hbm_peak) if k2b_ms else None,
                "traffic": ncu_traffic("k_explain_pass<1>", wl), "traffic_source": "profiles/traffic.json (ncu --set full, same workload)",
                "algorithmic_bytes": int(k2b_bytes), "peak_source": peak_src, "dominant_launch": dominant,
                "k2a": {"kernel": "k_classify (validity + singleton of every peak x breakage pair)", "algorithmic_bytes": int(k2a_bytes), "ms": k2a_ms,
                        "achieved": k2a_bytes / (k2a_ms * 1e-3) / 1e9 if k2a_ms else None,
                        "frac": (k2a_bytes / (k2a_ms * 1e-3) / 1e9 / hbm_peak) if k2a_ms else None, "traffic": ncu_traffic("k_classify", wl)}}
    launches = sum(v[1] for v in stats.values())
    ph = ctx.explain_phase_ns().astype(np.int64)
    ph = ph[ph > 0]
    phase_us = [round(float(x) * 1e-3, 2) for x in np.diff(ph)] if len(ph) > 1 else None  # see sst_explain_phase_ns

    # ---- the same step on a batch `large_factor` times larger (rank 0): where the latency of the ~19 dependent
    # phases of the pass is amortised and bytes per second mean something
    large = None
    if rank == 0 and args.large_factor > 1:
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
                 "classify_gbs": k2a_bytes * f / (cls_ms * 1e-3) / 1e9, "classify_frac": k2a_bytes * f / (cls_ms * 1e-3) / 1e9 / hbm_peak}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": workload_config(wl, args),
        "compositions_per_sec": comps_all * args.steps / (total_ms * 1e-3), "compositions_per_step": comps_all,
        "roots_per_step_rank0": int(n_roots),
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline, "kernels": kernels, "pass_phase_us": phase_us, "large_batch": large, "table_build": table_info,
    }

    if rank == 0 and not args.no_parity:
        line["parity"] = parity_gate(dp, wl, batch, valid, dev)
    if rank == 0 and not args.no_cpu_baseline:
        rows = table_rows_for(wl)
        workers = os.cpu_count() or 1
        rate0, _, _, _ = cpu_leg(wl, rows, max(50, min(400, wl.n_peaks)), False, workers)
        n_sample = int(max(100, min(wl.n_peaks, rate0 * 15.0)))
        rate, comps, dt, n_e = cpu_leg(wl, rows, n_sample, False, workers)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": workers, "kind": "port",
                                "sample": f"first {n_sample} of {wl.n_peaks} peaks ({n_sample * (len(wl.valid_mass) // wl.n_peaks)} validity + {n_e} explanation calls), {dt:.1f} s wall",
                                "compositions_per_sec": comps / dt,
                                "what": "oracle/oracle_py.py: pure-Python restatement of the reference functions (the reference is pure Python)"}
        n_c = int(min(wl.n_peaks, max(2000, n_sample * 20)))
        rate_c, comps_c, dt_c, n_ec = cpu_leg(wl, rows, n_c, True, workers)
        line["cpu_baseline_c"] = {"value": rate_c, "unit": UNIT, "cores": workers, "kind": "port",
                                  "sample": f"first {n_c} peaks, {dt_c:.1f} s wall", "what": "oracle/oracle.c (same algorithm in C)"}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def parity_gate(dp, wl, batch, valid, dev):
    """Same-run parity: table SHA-256 against the golden value, and a subsample of calls against the C oracle."""
    from oracle import oracle_c, oracle_py

    out = {}
    gold = json.loads((ROOT / "tests" / "golden" / "tables_sha.json").read_text())
    weights = [m.mass for m in dp.masses]
    if weights == gold["full"]["weights"] and dev.C == gold["full"]["shape"][1]:
        out["table_sha256_matches_reference"] = hashlib.sha256(dev.download().tobytes()).hexdigest() == gold["full"]["sha256"]
    tab = oracle_c.build_bit_table(weights, max(weights) * wl.max_seq_length, 32)
    if "table_sha256_matches_reference" not in out:
        out["table_equals_c_oracle"] = bool(np.array_equal(dev.download(), tab))
    rows = [oracle_py.Row(m.mass, m.is_modification, m.modification_rate) for m in dp.masses]
    ind = oracle_py.individual_budgets(rows, dp.seq.max_len)
    is_mod = [r.is_modification for r in rows]
    rng = np.random.default_rng(2)
    idx = rng.choice(len(wl.explain_mass), size=min(1000, len(wl.explain_mass)), replace=False)
    h = hashlib.sha256()
    ok = True
    for p in idx:
        tg, th = oracle_py.integerise(float(wl.explain_mass[p]), float(wl.explain_thr[p]), dp.precision, dp.tolerance)
        r, off, _ = oracle_c.explain(tab, 32, weights, is_mod, ind, tg, th, wl.max_modifications, True)
        want = sorted(tuple(int(x) for x in r[off[i]:off[i + 1]]) for i in range(len(off) - 1) if off[i + 1] > off[i])
        got = batch.canonical(int(p))
        ok &= got == want
        h.update(repr(got).encode())
    out["explain_subsample"] = len(idx)
    out["explain_ok"] = bool(ok)
    out["explain_digest"] = h.hexdigest()[:16]
    vi = rng.choice(len(wl.valid_mass), size=min(4000, len(wl.valid_mass)), replace=False)
    vok = True
    n_off = len(wl.breakage)
    flags = valid.flags  # [breakage, peak]; wl.valid_* are peak-major
    for p in vi:
        tg, th = oracle_py.integerise(float(wl.valid_mass[p]), float(wl.valid_thr[p]), dp.precision, dp.tolerance)
        try:
            want = 1 if oracle_c.is_valid(tab, 32, tg, th) else 0
        except NotImplementedError:
            want = 2
        vok &= int(flags[p % n_off, p // n_off] & 3) == want
    out["validity_subsample"] = len(vi)
    out["validity_ok"] = bool(vok)
    return out


if __name__ == "__main__":
    main()
