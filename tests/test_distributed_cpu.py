"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: source-window sharding, count exchange, list all-gather,
field-level dealing.  The per-rank slices come from the CPU oracle standing in for the GPU generator, so what is
checked is exactly the N>1 plumbing: rank-order concatenation of window slices == the serial list, bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import xgtest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _window_slice(full, tiles, lo, hi):
    """entries of the serial list whose source cell (concatenated index) lies in [lo, hi)"""
    offs = np.cumsum([0] + [a * b for a, b in tiles])[:-1]
    nx = np.array([a for a, b in tiles])
    s = offs[full["t_in"]] + full["j_in"].astype(np.int64) * nx[full["t_in"]] + full["i_in"]
    m = (s >= lo) & (s < hi)
    return {k: np.ascontiguousarray(v[m]) for k, v in full.items() if k != "nxgrid"}, s


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pkg = xgtest.package()
        from fre_nctools_b200 import distributed as D
        ni, nlon, nlat = 8, 36, 18
        lonc, latc = pkg.cubed_sphere_grid(ni)
        lon2, lat2 = pkg.latlon_grid(nlon, nlat)
        tiles = [(ni, ni)] * 6
        full = xgtest.oracle_setup(lonc, latc, lon2, lat2, 2)
        # candidate-pair counts per source cell stand-in: exchange cells per source cell (same partition rule)
        _, s = _window_slice(full, tiles, 0, 6 * ni * ni)
        cnt = np.bincount(s, minlength=6 * ni * ni)
        bounds = D.window_bounds_from_counts(cnt, world)
        assert bounds[0] == 0 and bounds[-1] == 6 * ni * ni and all(a <= b for a, b in zip(bounds, bounds[1:]))
        mine, _ = _window_slice(full, tiles, bounds[rank], bounds[rank + 1])
        n = mine["area"].size
        off, total, counts = D.exchange_offsets(n)
        assert total == full["nxgrid"] and counts[rank] == n
        assert abs(counts[0] - counts[1]) <= 2 * cnt.max()              # balanced split
        assert np.array_equal(full["area"][off:off + n], mine["area"])  # this rank's slice sits at its global offset
        g = D.allgather_xgrid({k: torch.from_numpy(v) for k, v in mine.items()})
        for k in D.INT_KEYS + D.F64_KEYS:
            assert np.array_equal(g[k].numpy(), full[k]), k             # rank-order concatenation == serial list
        lv = D.shard_field_levels(7)
        assert lv == list(range(rank, 7, world))
        # apply on the gathered list == apply on the serial list (oracle), for this rank's field-levels
        rng = np.random.default_rng(3)
        fields = rng.uniform(0, 1, (7, 6 * ni * ni))
        x = {k: v.numpy() for k, v in g.items()}
        for b in lv:
            a = xgtest.oracle_apply(x, 1, tiles, fields[b], nlon, nlat)
            w = xgtest.oracle_apply(full, 1, tiles, fields[b], nlon, nlat)
            assert np.array_equal(a, w)
        q.put((rank, "ok"))
    except Exception as e:                                              # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_window_sharding_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", f"rank {r}: {msg}"


def test_single_process_defaults():
    pkg = xgtest.package()
    from fre_nctools_b200 import distributed as D
    assert D.exchange_offsets(5) == (0, 5, [5])
    assert D.shard_field_levels(3) == [0, 1, 2]
    assert D.window_bounds_from_counts(np.array([1, 1, 1, 1]), 2) == [0, 2, 4]


def test_cost_balanced_shares():
    """rebalance_shares: a rank that measured a longer device time gets a smaller share of the candidate pairs; the host mirror
    of xgb_plan_partition_shares cuts the windows at the cumulative shares"""
    xgtest.package()
    from fre_nctools_b200 import distributed as D
    world, wpr = 4, 2
    shares = [1.0 / (world * wpr)] * (world * wpr)
    assert D.rebalance_shares(shares, [1.0, 1.0, 1.0, 1.0], world) == pytest.approx(shares)
    new = D.rebalance_shares(shares, [1.3, 1.0, 1.0, 1.2], world)
    assert sum(new) == pytest.approx(1.0)
    per_rank = [sum(new[r::world]) for r in range(world)]
    assert per_rank[0] < per_rank[3] < per_rank[1] == pytest.approx(per_rank[2])
    assert per_rank[0] / per_rank[1] == pytest.approx(1.0 / 1.3)
    # a fixed cost plus a cost per pair: repeated rounds converge to equal times
    fixed = [0.15, 0.0, 0.0, 0.1]
    s = shares
    for _ in range(6):
        t = [fixed[r] + 2.0 * sum(s[r::world]) for r in range(world)]
        s = D.rebalance_shares(s, t, world)
    t = [fixed[r] + 2.0 * sum(s[r::world]) for r in range(world)]
    assert max(t) / min(t) < 1.01
    # malformed input leaves the shares alone
    assert D.rebalance_shares(shares, [1.0, 0.0, 1.0, 1.0], world) == shares and D.rebalance_shares(shares, [1.0], world) == shares
    counts = np.array([4, 4, 4, 4, 4, 4, 4, 4])
    assert D.window_bounds_from_counts(counts, 2, [1.0, 1.0]) == D.window_bounds_from_counts(counts, 2) == [0, 4, 8]
    assert D.window_bounds_from_counts(counts, 2, [1.0, 3.0]) == [0, 2, 8]
    assert D.window_bounds_from_counts(counts, 4, [1.0, 1.0, 2.0, 4.0]) == [0, 1, 2, 4, 8]
