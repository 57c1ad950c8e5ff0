"""Multi-rank host logic on CPU: world_size-2 gloo, the C oracle standing in for the device call.

The partition / gather code is the product's (spectrseqtools_b200/sharding.py); only the per-rank compute
callback is replaced (no GPU in this container) — by the oracle, which is allowed in tests.
"""
import os
import socket
import sys
import types

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle_c, oracle_py
from spectrseqtools_b200 import sharding

WEIGHTS = [0, 1201, 1333, 1479, 1612, 2050]
IS_MOD = [False, False, False, True, False, True]
RATES = [0.0, 1.0, 1.0, 0.5, 1.0, 0.5]
MAX_LEN = 12
TOL = 2e-4
PREC = 1e-3


def _inputs():
    rng = np.random.default_rng(7)
    picks = rng.integers(1, len(WEIGHTS), size=(60, 6))
    n = rng.integers(0, 7, size=60)
    masses = np.array([sum(WEIGHTS[i] for i in row[:k]) for row, k in zip(picks, n)], dtype=np.float64) * PREC
    masses += rng.uniform(-2e-3, 2e-3, size=len(masses))
    masses[5] = 1e9  # far beyond the table -> out-of-table status
    thr = np.where(rng.random(len(masses)) < 0.5, np.nan, rng.uniform(1e-3, 6e-3, size=len(masses)))
    return masses, thr


def _oracle_local(table):
    rows = [oracle_py.Row(w, m, r) for w, m, r in zip(WEIGHTS, IS_MOD, RATES)]
    ind = oracle_py.individual_budgets(rows, MAX_LEN)

    def fn(m, t, k):
        status, counts, recs = [], [], []
        for mass, th in zip(m, t):
            tg, it = oracle_py.integerise(float(mass), None if np.isnan(th) else float(th), PREC, TOL)
            try:
                r, off, zero = oracle_c.explain(table, 32, WEIGHTS, IS_MOD, ind, tg, it, k, True)
            except NotImplementedError:
                status.append(2)
                counts.append(0)
                continue
            lists = [tuple(int(x) for x in r[off[i]:off[i + 1]]) for i in range(len(off) - 1) if off[i + 1] > off[i]]
            status.append(1 if (tg - it <= 0 <= tg + it) else 0)
            counts.append(len(lists))
            for c in lists:
                rec = np.zeros(8, dtype=np.uint8)
                rec[: len(c)] = sorted(c)  # row indices
                recs.append(rec)
        recs = np.stack(recs) if recs else np.zeros((0, 8), dtype=np.uint8)
        return np.array(status, np.uint8), np.array(counts, np.int64), recs

    return fn


def _dp_stub():
    masses = [types.SimpleNamespace(mass=w, names=[f"n{w}"], is_modification=m, modification_rate=r)
              for w, m, r in zip(WEIGHTS, IS_MOD, RATES)]
    return types.SimpleNamespace(masses=masses)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        table = oracle_c.build_bit_table(WEIGHTS, max(WEIGHTS) * 35, 32)
        masses, thr = _inputs()
        batch = sharding.explain_masses_sharded(masses, _dp_stub(), max_modifications=3, thresholds=thr,
                                                local_fn=_oracle_local(table))
        valid = sharding.are_valid_masses_sharded(
            masses, _dp_stub(), thr,
            local_fn=lambda m, t: np.array([_valid_code(table, float(a), b) for a, b in zip(m, t)], np.uint8))
        if rank == 0:
            q.put((batch.status, batch.offsets, batch.records, valid))
        else:
            assert batch is None and valid is None
            q.put("ok")
    finally:
        dist.destroy_process_group()


def _valid_code(table, mass, th):
    tg, it = oracle_py.integerise(mass, None if np.isnan(th) else float(th), PREC, TOL)
    try:
        return 1 if oracle_c.is_valid(table, 32, tg, it) else 0
    except NotImplementedError:
        return 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_partition_covers_everything_and_balances_mass_bins():
    masses, thr = _inputs()
    for world in (1, 2, 3, 8):
        parts = sharding.partition_by_mass(masses, thr, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(masses)))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1
        for p in parts:
            assert np.all(np.diff(p) > 0)
    # mass-bin interleave: every rank sees the heavy end
    m = np.arange(1000, dtype=np.float64)
    parts = sharding.partition_by_mass(m, None, 4)
    assert all(p.max() >= 996 for p in parts)


def test_two_rank_gloo_gather_equals_single_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    status, offsets, records, valid = next(g for g in got if not isinstance(g, str))

    table = oracle_c.build_bit_table(WEIGHTS, max(WEIGHTS) * 35, 32)
    masses, thr = _inputs()
    st1, cnt1, rec1 = _oracle_local(table)(masses, thr, 3)
    assert np.array_equal(status, st1)
    assert np.array_equal(np.diff(offsets), cnt1)
    assert np.array_equal(records, rec1)  # same order: per-peak grouping and in-peak order survive the gather
    assert status[5] == 2 and cnt1.sum() > 50
    want_valid = np.array([_valid_code(table, float(a), b) for a, b in zip(masses, thr)], np.uint8)
    assert np.array_equal(valid, want_valid)


# ---------------------------------------------------------------- GPU: the real device callback on two ranks
def _gpu_workload():
    from spectrseqtools_b200 import mass_table as MT
    from spectrseqtools_b200 import synthetic as S

    wl = S.make_workload("C2", 3000)
    seq = MT.SequenceInformation(max_len=wl.max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
    dp = MT.DynamicProgrammingTable(S.alphabet_frame(wl.alphabet), 32, wl.ppm, 1e-3, seq, device=0)
    masses = np.concatenate([wl.explain_mass, [1e9, 0.0002]])       # + an out-of-table mass and a zero-window mass
    thr = np.concatenate([wl.explain_thr, [0.01, np.nan]])
    return wl, dp, masses, thr


def _gpu_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl, dp, masses, thr = _gpu_workload()
        batch = sharding.explain_masses_sharded(masses, dp, max_modifications=wl.max_modifications, thresholds=thr)
        valid = sharding.are_valid_masses_sharded(wl.valid_mass[:5000], dp, wl.valid_thr[:5000])
        if rank == 0:
            q.put((batch.status, batch.offsets, batch.records, valid))
        else:
            assert batch is None and valid is None
            q.put("ok")
    finally:
        dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_device_path_equals_single_rank(world):
    """explain_masses_sharded / are_valid_masses_sharded with the CUDA callback on `world` ranks (one process per
    rank, all on GPU 0 here; one GPU each under torchrun) == the single-rank batch call: status, offsets, records.
    world = 3 leaves ragged shards; the appended out-of-table and zero-window masses land on different ranks."""
    from spectrseqtools_b200 import mass_explanation as ME

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    status, offsets, records, valid = next(g for g in got if not isinstance(g, str))
    wl, dp, masses, thr = _gpu_workload()
    one = ME.explain_masses(masses, dp, max_modifications=wl.max_modifications, thresholds=thr)
    assert np.array_equal(status, one.status) and status[-2] == 2 and status[-1] & 1
    assert np.array_equal(offsets, one.offsets) and offsets[-1] > 2000
    W = min(records.shape[1], one.records.shape[1])
    assert np.array_equal(records[:, :W], one.records[:, :W]) and not records[:, W:].any() and not one.records[:, W:].any()
    assert np.array_equal(valid, ME.are_valid_masses(wl.valid_mass[:5000], dp, wl.valid_thr[:5000]))


if __name__ == "__main__":
    sys.exit(0)


def _shm_worker(rank, world, tag, steps, q):
    """One rank of the shared-memory gather: publishes `steps` consecutive steps, rank 0 collects and checks them."""
    import numpy as np

    from spectrseqtools_b200 import sharding

    g = sharding.ShmGather(tag, rank, world, capacity=1 << 20)
    try:
        ok = True
        for seq in range(1, steps + 1):
            a = np.full(100 + rank, seq * 10 + rank, dtype=np.uint32)
            b = np.arange(6, dtype=np.uint8).reshape(3, 2) + seq + rank
            g.publish(seq, [a, b])
            if rank == 0:
                got = g.collect(seq)
                for r, (ga, gb) in enumerate(got):
                    ok &= ga.shape == (100 + r,) and bool((ga == seq * 10 + r).all())
                    ok &= gb.shape == (3, 2) and bool((gb == (np.arange(6, dtype=np.uint8).reshape(3, 2) + seq + r)).all())
        if rank == 0:
            g.release(steps)
            q.put(ok)
        else:  # stay until rank 0 has read the last step
            sharding.ShmGather._spin(lambda: g._hdr[1] < steps, "last acknowledgement", timeout=30.0)
            q.put(True)
    finally:
        g.close()


def test_shm_gather_many_consecutive_steps():
    """The gather the strong-scaling bench uses: consecutive step numbers, a segment half reused every second step, the
    publisher waits for the release of step seq - 2 (a gap in the numbering used to hang it for ever, and an
    acknowledgement sent before the views were read let the publisher overwrite them)."""
    import multiprocessing as mp
    import os

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    tag = f"sstb200_test_{os.getpid()}"
    procs = [ctx.Process(target=_shm_worker, args=(r, 2, tag, 40, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(res)
