"""Two-GPU test (NCCL): every rank generates its source-cell window, counts and slices are exchanged, and the
rank-order concatenation must equal the oracle's serial list bit for bit; field-levels dealt to ranks remap to the
oracle's values.  Skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import xgtest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pkg = xgtest.package()
        from fre_nctools_b200 import distributed as D
        ni, nlon, nlat = 24, 144, 72
        lonc, latc = pkg.cubed_sphere_grid(ni)
        lon2, lat2 = pkg.latlon_grid(nlon, nlat)
        tiles = [(ni, ni)] * 6
        plan = pkg.XgridPlan(rank)
        plan.set_dst(lon2, lat2)
        plan.set_src(lonc, latc)
        bounds = plan.partition(world)
        plan.set_src_window(bounds[rank], bounds[rank + 1])
        n = plan.generate(pkg.CONSERVE_ORDER2)
        dev = torch.device("cuda", rank)
        off, total, counts = D.exchange_offsets(n, dev)
        full = xgtest.oracle_setup(lonc, latc, lon2, lat2, 2)
        assert total == full["nxgrid"], (total, full["nxgrid"])
        plan.sync()
        loc = plan.result_device()
        g = D.allgather_xgrid({k: loc[k] for k in D.INT_KEYS + D.F64_KEYS})
        torch.cuda.synchronize()
        for k in D.INT_KEYS + D.F64_KEYS:
            assert np.array_equal(g[k].cpu().numpy(), full[k]), k
        # every rank applies its share of the field-levels with the whole list
        p2 = pkg.XgridPlan(rank)
        p2.set_xgrid(tiles, nlon, nlat, g)
        rng = np.random.default_rng(11)
        fields = rng.uniform(0, 1, (5, 6 * ni * ni))
        lv = D.shard_field_levels(5)
        out = p2.apply(1, np.ascontiguousarray(fields[lv]).reshape(-1), len(lv)).reshape(len(lv), -1)
        for o, b in zip(out, lv):
            assert np.array_equal(o, xgtest.oracle_apply(full, 1, tiles, fields[b], nlon, nlat)), b
        q.put((rank, "ok"))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_gpu_windows_concatenate_to_serial_list():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", f"rank {r}: {msg}"
