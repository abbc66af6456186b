#!/usr/bin/env python
"""bench.py — exchange-grid weight generation throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c768] [--impl reference]

A "step" is one complete weight generation (candidate search, clip, area/centroids, ordered
compaction, order-2 centroid correction) of the workload, sharded by source-cell windows over the
N ranks.  `value` is whole-job exchange cells per second with the grids already resident in HBM;
`e2e` is the same job through the C ABI with HOST buffers (grid upload and result download inside
the timed region).  N > 1 is launched by torchrun (one rank per GPU, NCCL).

--impl reference times the reference's own CPU implementation (oracle/_ref, the unmodified
FRE-NCtools sources compiled by oracle/Makefile; the oracle port if that was never built) on all
host cores, each step a bounded sample of destination row bands of the same workload.
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "xgrid_cells_per_sec"
UNIT = "xgrid cells/s"

# name -> (cubed-sphere N, nlon, nlat, order)
WORKLOADS = {
    "c768": (768, 2880, 1440, 2),     # BASELINE.json configs[3]: C768 -> 1/8 degree, order 2 (the metric's config)
    "c96": (96, 1440, 720, 2),        # configs[1] weights
    "c48": (48, 360, 180, 1),         # configs[0]
    "c384": (384, 1440, 720, 2),
    "c3072": (3072, 11520, 5760, 2),  # configs[4] (sizing run)
    "c3072o1": (3072, 11520, 5760, 1),  # the reference's one published figure (conserve_order1, 2558 s on 641 ranks)
}

# Algorithmic FP64 operations per emitted exchange cell (source-level + - * / compare fabs = 1, sin/cos = 1),
# counted by the op-counting build of the oracle (oracle/count_ops.py, see DESIGN.md "Algorithmic work").
OPS_PER_XCELL = {1: 4.6e2, 2: 6.5e2}
# bytes the clip kernel must move per candidate pair (order 2): pair 8 + src/dst cell polygons 2*(8*8+1) ... see DESIGN.md
CLIP_BYTES_PER_PAIR = {1: 8 + 8, 2: 8 + 24}


def workload_label(name):
    n, nlon, nlat, order = WORKLOADS[name]
    return f"C{n} gnomonic_ed cubed sphere (6x{n}x{n}) -> {nlon}x{nlat} lat-lon, conserve_order{order} weight generation"


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU reference / oracle sample (bounded): destination row bands on all host cores
# ---------------------------------------------------------------------------------------------
_G = {}


def _band_worker(args):
    import xgtest
    jsc, jec, order, use_ref, keep = args
    lonc, latc, lon2, lat2 = _G["grids"]
    t0 = time.perf_counter()
    if use_ref:
        r = xgtest.ref_setup(lonc, latc, lon2, lat2, order, jsc=jsc, jec=jec)
    else:
        r = xgtest.oracle_setup(lonc, latc, lon2[jsc:jec + 2], lat2[jsc:jec + 2], order)
    dt = time.perf_counter() - t0
    lists = None
    if keep:                      # the band's list, rows made global, for the parity check against the GPU's list
        lists = {k: r[k] for k in ("t_in", "i_in", "j_in", "i_out", "area")}
        lists["j_out"] = r["j_out"] + jsc
    return r["nxgrid"], dt, lists


def cpu_sample(name, rows_per_worker, cores, pool=None, keep=False):
    """One bounded sample: `cores` destination bands of rows_per_worker rows, evenly spread over the
    latitudes, each run by one process exactly as a fregrid_parallel rank owning that band would
    (fregrid_util.c:592-603 layout {1,npes}).  Returns (xcells, wall seconds, kind[, per-band lists])."""
    import xgtest
    n, nlon, nlat, order = WORKLOADS[name]
    use_ref = xgtest.ref_lib() is not None
    rows_per_worker = max(1, min(rows_per_worker, nlat // cores))
    starts = [int((k + 0.5) * nlat / cores) for k in range(cores)]
    jobs = [(min(s, nlat - rows_per_worker), min(s, nlat - rows_per_worker) + rows_per_worker - 1, order, use_ref, keep) for s in starts]
    t0 = time.perf_counter()
    res = pool.map(_band_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    out = (sum(r[0] for r in res), wall, ("reference" if use_ref else "port"))
    return out + ([(j[0], j[1], r[2]) for j, r in zip(jobs, res)],) if keep else out


def parity_check_bands(gpu, bands):
    """the GPU's exchange-grid list (host dict) restricted to each band's destination rows must be the band's list the CPU
    reference produced: same cells in the same order, areas bit-identical when the host libm is the reference's (else 1e-9).
    tile1_distance of a band run is not comparable (its per-source-cell sums only see the band).  -> cells compared"""
    import xgtest
    exact = xgtest.libm_matches_ref_trig()
    n = 0
    for jsc, jec, b in bands:
        m = (gpu["j_out"] >= jsc) & (gpu["j_out"] <= jec)
        if int(m.sum()) != b["area"].size:
            raise SystemExit(f"bench.py: PARITY FAILURE rows {jsc}..{jec}: GPU {int(m.sum())} cells, CPU reference {b['area'].size}")
        for k in ("t_in", "i_in", "j_in", "i_out", "j_out"):
            if not np.array_equal(gpu[k][m], b[k]):
                raise SystemExit(f"bench.py: PARITY FAILURE rows {jsc}..{jec}: {k} differs from the CPU reference")
        ok = np.array_equal(gpu["area"][m], b["area"]) if exact else bool(np.max(np.abs(gpu["area"][m] - b["area"]) / b["area"]) < 1e-9)
        if not ok:
            raise SystemExit(f"bench.py: PARITY FAILURE rows {jsc}..{jec}: xgrid_area differs from the CPU reference")
        n += b["area"].size
    return n


def make_pool(name, cores):
    """the CPU legs build their grids WITHOUT the product library: the reference's own cubed-sphere generator (oracle/_ref) and
    get_output_grid_by_size restated in numpy (xgtest.latlon_grid_np).  Only when oracle/_ref was never built (then the CPU leg
    runs the oracle port) does the source grid come from the product's host-side generator."""
    import xgtest
    n, nlon, nlat, order = WORKLOADS[name]
    if xgtest.ref_lib() is not None:
        lonc, latc = xgtest.ref_cubed_sphere(n)           # the reference's own generator
    else:
        lonc, latc = xgtest.package().cubed_sphere_grid(n)
    lon2, lat2 = xgtest.latlon_grid_np(nlon, nlat)
    _G["grids"] = (lonc, latc, lon2, lat2)
    ctx = mp.get_context("fork")                          # grids are inherited copy-on-write
    return ctx.Pool(cores)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    cores = os.cpu_count() or 1
    pool = make_pool(name, cores)
    # calibrate one row per worker, then size the per-step sample to finish the whole run in a few minutes
    x0, t0, kind = cpu_sample(name, 1, cores, pool)
    budget = min(25.0, max(t0, 150.0 / (args.steps + args.warmup)))
    rows = max(1, min(int(budget / max(t0, 1e-3)), WORKLOADS[name][2] // cores))
    for _ in range(max(0, args.warmup - 1)):
        cpu_sample(name, rows, cores, pool)
    tot_x, tot_t = 0, 0.0
    for _ in range(args.steps):
        x, t, kind = cpu_sample(name, rows, cores, pool)
        tot_x += x; tot_t += t
    pool.close()
    val = tot_x / tot_t
    n, nlon, nlat, order = WORKLOADS[name]
    sample = (f"{cores} destination row bands x {rows} rows of {nlat} (evenly spaced latitudes), all 6 source tiles, "
              f"one process per band; {tot_x // max(args.steps, 1)} xcells per step")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_label(name), "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# apply leg (BASELINE.json metric, second half: "apply GB/s"): configs[1] = C96 -> 1440x720 order 2, 33 levels x 12 times
# ---------------------------------------------------------------------------------------------
APPLY_CFG = {"ni": 96, "nlon": 1440, "nlat": 720, "levels": 33, "times": 12}


def apply_leg(pkg, torch, dist, rank, world, local, steps, warmup):
    """grad_c2l + do_scalar_conserve_interp for all 396 field-levels of configs[1] per step.  Field-levels are dealt
    round-robin to the ranks (no data-path collective: every rank holds the whole exchange grid, 53 MB).
    Algorithmic bytes per step: nxgrid*32 once + B*(Nsrc*8 data + Nsrc*16 gradients written + Nsrc*16
    gradients read + Ndst*8 output)."""
    import xgtest
    ni, nlon, nlat = APPLY_CFG["ni"], APPLY_CFG["nlon"], APPLY_CFG["nlat"]
    B = APPLY_CFG["levels"] * APPLY_CFG["times"]
    mine = list(range(rank, B, world))
    nb = len(mine)
    lonc, latc, lont, latt = pkg.cubed_sphere_grid(ni, centers=True)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    hm = xgtest.cubed_sphere_halo_map(lonc, latc)
    plan = pkg.XgridPlan(local)
    plan.set_dst(lon2, lat2)
    plan.set_src(lonc, latc)
    nx = plan.generate(pkg.CONSERVE_ORDER2)
    plan.apply_setup()
    plan.grad_setup(xgtest.with_halo(lont.reshape(-1), hm), xgtest.with_halo(latt.reshape(-1), hm))
    ncell, nhalo, ndst = 6 * ni * ni, 6 * (ni + 2) ** 2, nlon * nlat
    rng = np.random.default_rng(1234 + rank)
    base = xgtest.smooth_field(lont, latt)
    f = np.stack([base + 0.01 * (b % 33) + 0.001 * (b // 33) if b % 2 == 0 else rng.uniform(0, 1, ncell) for b in mine])
    h_in = torch.from_numpy(xgtest.with_halo(f, hm).reshape(-1)).pin_memory()
    h_out = torch.empty(nb * ndst, dtype=torch.float64).pin_memory()
    dev = torch.device("cuda", local)
    d_in = h_in.to(dev)
    d_out = torch.empty(nb * ndst, dtype=torch.float64, device=dev)
    ext = torch.cuda.ExternalStream(plan.stream, device=dev)
    opcode = pkg.CONSERVE_ORDER2
    launches0 = pkg.kernel_launches()

    def timed(fn, n):
        for _ in range(max(warmup, 1)):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
        for _ in range(n):
            fn()
        with torch.cuda.stream(ext):
            e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms = timed(lambda: plan.regrid(opcode, d_in, nb, out=d_out), max(steps, 3))
    launches = (pkg.kernel_launches() - launches0) // (max(steps, 3) + max(warmup, 1))
    e2e_ms = timed(lambda: plan.regrid(opcode, h_in.numpy(), nb, out=h_out.numpy()), 3)
    plan.close()
    bytes_step = nx * 32 * world + B * (ncell * 8 + ncell * 16 * 2 + ndst * 8)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    gbs = bytes_step / (ms * 1e-3) * 1e-9             # whole job, all ranks
    gbs_gpu = gbs / world                              # what one GPU's memory system delivers: the roofline compares THIS with one GPU's peak
    return {"workload": f"C{ni} -> {nlon}x{nlat} conserve_order2 remap (grad_c2l + apply) of {B} field-levels "
                        f"({APPLY_CFG['levels']} levels x {APPLY_CFG['times']} times), nxgrid {nx}",
            "metric": "apply_GB_per_sec", "value": gbs, "unit": "GB/s (algorithmic bytes)", "ms_per_step": ms,
            "field_levels_per_sec": B / (ms * 1e-3), "algorithmic_bytes_per_step": int(bytes_step), "gpu_launches_per_step": int(launches),
            "roofline": {"bound": "hbm", "achieved": gbs_gpu, "peak": hbm_peak, "unit": "GB/s per GPU", "frac": gbs_gpu / hbm_peak,
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s", "kernel": ("apply_packed_kernel + grad_c2l_kernel" if os.environ.get("XGB_APPLY_PACKED") == "1" else "apply_rec_kernel + grad_c2l_rec_kernel"),
                         "note": "per GPU: whole-job algorithmic GB/s divided by the number of ranks, against ONE GPU's measured HBM peak"},
            "e2e": {"value": bytes_step / (e2e_ms * 1e-3) * 1e-9, "unit": "GB/s (algorithmic bytes)", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h_in.numel() * 8 * world), "d2h_bytes_per_step": int(h_out.numel() * 8 * world)},
            "sharding": f"{B} field-levels dealt round-robin to {world} rank(s); every rank holds the whole exchange grid"}


# ---------------------------------------------------------------------------------------------
# great-circle leg: BASELINE configs[2] = 1/4 degree tripolar ocean grid (1440x1080) -> 1 degree lat-lon, create_xgrid_great_circle
# ---------------------------------------------------------------------------------------------
def _gc_row_worker(args):
    import xgtest
    j0, nrows = args
    tl, ta, lon2, lat2 = _G["gc"]
    t0 = time.perf_counter()
    r = xgtest.ref_setup([tl[j0:j0 + nrows + 1]], [ta[j0:j0 + nrows + 1]], lon2, lat2, 1 | xgtest.GREAT_CIRCLE)
    return j0, nrows, r, time.perf_counter() - t0


def gc_leg(pkg, torch, local, steps, warmup, with_cpu):
    """device-resident great-circle weight generation of configs[2], CUDA events on the plan's stream; the CPU baseline is the
    unmodified reference (setup_conserve_interp with GREAT_CIRCLE) on one source row per host core — the reference scans every
    destination cell for every source cell, so a row is a fair sample of its cost — and doubles as the parity check."""
    import xgtest
    if xgtest.ref_lib() is None:
        return None                                      # the tripolar generator is the reference's own (oracle/_ref)
    tl, ta = xgtest.tripolar_grid(2880, 2160)
    lon2, lat2 = xgtest.latlon_grid_np(360, 180)
    plan = pkg.XgridPlan(local)
    plan.set_dst(lon2, lat2); plan.set_src([tl], [ta])
    op = pkg.CONSERVE_ORDER1 | pkg.GREAT_CIRCLE
    n = plan.generate(op)
    for _ in range(max(warmup, 1)):
        plan.generate(op)
    ext = torch.cuda.ExternalStream(plan.stream, device=torch.device("cuda", local))
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    k = max(steps, 3)
    with torch.cuda.stream(ext):
        e0.record()
    for _ in range(k):
        plan.generate(op)
    with torch.cuda.stream(ext):
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    out = {"workload": "1/4 degree tripolar ocean grid (make_hgrid tripolar, 1440x1080 cells) -> 360x180 lat-lon, create_xgrid_great_circle (order 1)",
           "metric": METRIC, "value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "nxgrid": int(n), "candidate_pairs": int(plan.npairs)}
    if with_cpu:
        x = plan.result_host()
        cores = os.cpu_count() or 1
        ny1 = tl.shape[0] - 1
        rows = sorted(set(int((q + 0.5) * ny1 / cores) for q in range(cores)))
        _G["gc"] = (tl, ta, lon2, lat2)
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_gc_row_worker, [(j, 1) for j in rows], chunksize=1)
        wall = time.perf_counter() - t0
        checked = 0
        for j0, nrows, r, _ in res:
            m = (x["j_in"] >= j0) & (x["j_in"] < j0 + nrows)
            same = int(m.sum()) == r["nxgrid"] and all(np.array_equal(x[kk][m], r[kk] + (j0 if kk == "j_in" else 0)) for kk in ("i_in", "j_in", "i_out", "j_out"))
            if not same or (r["nxgrid"] and np.max(np.abs(x["area"][m] - r["area"])) / 6371000.0 ** 2 > 8e-15):
                raise SystemExit(f"bench.py: PARITY FAILURE (great circle) source row {j0}")
            checked += r["nxgrid"]
        tot = sum(r["nxgrid"] for _, _, r, _ in res)
        out["cpu_baseline"] = {"value": tot / wall, "unit": UNIT, "cores": cores, "kind": "reference",
                               "sample": f"{len(rows)} source rows of {ny1} (evenly spaced), one process per row, all 64800 destination cells: {tot} xcells in {wall:.1f} s"}
        out["parity_checked_xcells"] = checked
    plan.close()
    return out


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def list_checksum(torch, r, n, nlon, order):
    """order-independent 64-bit sums over a device-resident exchange-grid list (plan.result_device()): count, cell-pair keys,
    bit patterns of xgrid_area and tile1_distance.  The pieces of an N-GPU run must add up to the single-GPU list's sums."""
    s = (r["t_in"].to(torch.int64) * (n * n) + r["j_in"].to(torch.int64) * n + r["i_in"].to(torch.int64))
    d = r["j_out"].to(torch.int64) * nlon + r["i_out"].to(torch.int64)
    out = [torch.tensor(s.numel(), dtype=torch.int64, device=s.device), (s * 1315423911 + d * 2654435761).sum(),
           r["area"].view(torch.int64).sum()]
    if order == 2:
        out += [r["di"].view(torch.int64).sum(), r["dj"].view(torch.int64).sum()]
    return torch.stack(out)


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = ge.load_package()
    name = args.workload
    n, nlon, nlat, order = WORKLOADS[name]
    opcode = pkg.CONSERVE_ORDER2 if order == 2 else pkg.CONSERVE_ORDER1

    lonc, latc = pkg.cubed_sphere_grid(n)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    plan = pkg.XgridPlan(local)
    plan.set_dst(lon2, lat2)
    plan.set_src(lonc, latc)
    # Sharding (static, part of the plan): the source cells are cut into world*WPR contiguous windows of equal
    # candidate-pair count and dealt round-robin, so every rank holds polar and mid-latitude pieces alike (a pair near a
    # pole costs ~30 % more to clip).  Windows are contiguous pieces of the reference's emission order; the global list
    # is the windows' pieces in window order.
    WPR = 8 if world > 1 else 1
    bounds = plan.partition(world * WPR)
    my_windows = [(bounds[w], bounds[w + 1]) for w in range(rank, world * WPR, world)]

    def set_windows():
        if world > 1:
            plan.set_src_windows(my_windows)
        else:
            plan.set_src_window(bounds[0], bounds[1])
    set_windows()
    ext = torch.cuda.ExternalStream(plan.stream, device=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    counts = torch.zeros(world * WPR, dtype=torch.int64, device=dev)
    h_mine = torch.zeros(WPR, dtype=torch.int64).pin_memory()
    mine = torch.zeros(WPR, dtype=torch.int64, device=dev)

    # An asynchronous step enqueues one generate on the plan's stream (xgb_plan_generate_async) and, for N > 1, the path's one exchange behind
    # it: per-window counts taken on the device, all-gathered over NCCL -> global offsets of every piece.  Nothing waits for the
    # host inside the timed region; the count comes back with generate_finish() after it.  Default for N > 1 (2.42 -> 2.32 ms at
    # N = 2); at N = 1 the blocking xgb_plan_generate (one host synchronisation per step) measures the same and stays the default.
    use_async = args.async_steps or (world > 1 and not args.sync_steps)
    plan.generate(opcode)                           # sizes the buffers for this rank's windows (untimed)

    # ---- N > 1: cost-balanced windows (set-up, untimed).  Equal pair counts are not equal times: the ranks that own a pole spend
    # longer per pair (pole-cap enumeration by one warp, the pole cell's sequential sums).  Every rank measures the device time
    # of its share, the times are all-gathered, and the windows are re-cut with each rank's share of the pairs scaled by
    # mean(t) / t_rank (distributed.rebalance_shares -> xgb_plan_partition_shares); a few rounds settle it.  The plan that
    # results is static: the timed region below runs it unchanged, and the parity check after it covers the re-cut windows.
    balance = None
    if world > 1 and args.balance_rounds > 0:
        from importlib import import_module
        dmod = import_module(pkg.__name__ + ".distributed")
        shares = [1.0 / (world * WPR)] * (world * WPR)
        hist = []
        for _ in range(args.balance_rounds):
            plan.reset_phase_ms()
            for _k in range(3):
                plan.generate(opcode)
            _, acc, ngen = plan.phase_ms()
            tt = torch.tensor([sum(acc.values()) / max(ngen, 1)], dtype=torch.float64, device=dev)
            allt = torch.empty(world, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allt, tt)
            rank_ms = [float(v) for v in allt.tolist()]
            hist.append([round(v, 4) for v in rank_ms])
            if max(rank_ms) <= 1.02 * min(rank_ms):
                break
            shares = dmod.rebalance_shares(shares, rank_ms, world)
            bounds = plan.partition(world * WPR, shares)
            my_windows = [(bounds[w], bounds[w + 1]) for w in range(rank, world * WPR, world)]
            set_windows()
            plan.generate(opcode)                   # sizes the buffers for the new windows
        per_rank_share = [sum(shares[r::world]) for r in range(world)]
        balance = {"rounds": len(hist), "rank_device_ms_per_round": hist, "pair_share_per_rank": [round(v, 4) for v in per_rank_share],
                   "how": "windows re-cut by xgb_plan_partition_shares from the ranks' measured device time (set-up, untimed)"}

    def step():
        if use_async:
            plan.generate_async(opcode)
            if world > 1:
                plan.window_counts_device(mine)
                with torch.cuda.stream(ext):
                    dist.all_gather_into_tensor(counts, mine)
            return None
        nx = plan.generate(opcode)
        if world > 1:                           # the path's one exchange: per-window counts -> global offsets of every piece
            h_mine.copy_(torch.tensor(plan.window_counts(), dtype=torch.int64))
            with torch.cuda.stream(ext):
                mine.copy_(h_mine, non_blocking=True)
                dist.all_gather_into_tensor(counts, mine)
        return nx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    if use_async:
        plan.generate_finish()
    plan.reset_phase_ms()
    launches0 = pkg.kernel_launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        e0.record()
    nx = 0
    for _ in range(args.steps):
        nx = step()
    with torch.cuda.stream(ext):
        e1.record()
    barrier()
    if use_async:
        nx = plan.generate_finish()                 # the stream has drained: checks the kernels' error word, returns the count
        if world > 1:
            assert int(counts.sum().item()) > 0 and plan.window_counts() == mine.tolist()
    if rank == 0 and len(sampler.lines) < 5:
        # the timed region is a few tens of milliseconds, shorter than nvidia-smi's first answer on a big box: keep the SAME step
        # running (untimed) until the sampler has seen the clocks under this load
        t_end = time.perf_counter() + 0.6
        while time.perf_counter() < t_end and len(sampler.lines) < 5:
            if use_async:
                plan.generate_async(opcode); plan.generate_finish()
            else:
                plan.generate(opcode)
    clocks = sampler.stop() if rank == 0 else None
    # ---- N > 1: the ranks' pieces against the single-GPU list (untimed).  Every rank sums its piece; rank 0 then generates the
    # WHOLE problem on its own GPU (one window) and the all-reduced sums of the pieces must equal its sums exactly: same cells,
    # same areas, same tile1_distance, nothing lost or duplicated at the window boundaries.
    multi_parity = None
    if world > 1:
        mine_sum = list_checksum(torch, plan.result_device(), n, nlon, order)
        dist.all_reduce(mine_sum, op=dist.ReduceOp.SUM)
        whole = torch.zeros_like(mine_sum)
        if rank == 0:
            plan.set_src_window(0, plan.ncell_src)
            plan.generate(opcode)
            whole = list_checksum(torch, plan.result_device(), n, nlon, order)
            set_windows()
            plan.generate(opcode)
        dist.broadcast(whole, src=0)
        if not torch.equal(whole, mine_sum):
            raise SystemExit(f"bench.py: PARITY FAILURE: the {world} ranks' pieces do not add up to the single-GPU list "
                             f"(pieces {mine_sum.tolist()} vs whole {whole.tolist()})")
        multi_parity = {"checked_xcells": int(whole[0].item()), "against": "single-GPU list on rank 0: count, cell-pair keys, "
                        "xgrid_area and tile1_distance bit patterns (64-bit sums)", "equal": True}
    ms = e0.elapsed_time(e1) / args.steps
    launches = pkg.kernel_launches() - launches0
    _, phase_sum, ngen = plan.phase_ms()
    npairs = plan.npairs

    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([nx, npairs, launches], dtype=torch.int64, device=dev)
    clip_ms = torch.tensor([phase_sum["clip"] / max(ngen, 1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(clip_ms, op=dist.ReduceOp.MAX)
    per_rank = None
    if world > 1:                                   # per-rank step time and phase sums, for load-balance diagnosis
        mine_v = torch.tensor([e0.elapsed_time(e1) / args.steps, nx, npairs] + [phase_sum[k] / max(ngen, 1) for k in plan.PHASES],
                              dtype=torch.float64, device=dev)
        allv = torch.empty(world * mine_v.numel(), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allv, mine_v)
        per_rank = [dict(zip(["ms", "nxgrid", "pairs"] + list(plan.PHASES), row)) for row in allv.view(world, -1).tolist()]
    ms = float(t.item())
    nx_total, npairs_total, launches_total = (int(v) for v in tot.tolist())
    value = nx_total / (ms * 1e-3)

    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                              "ms_per_step": ms, "config": {"workload": workload_label(name), "nxgrid": nx_total, "candidate_pairs": npairs_total},
                              "phase_ms": {kk: vv / max(ngen, 1) for kk, vv in phase_sum.items()}, "note": "sizing run: value only"}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- end to end through the C ABI with host buffers (pinned), every step: H2D grids, generate, D2H result
    h_lon1 = torch.from_numpy(np.ascontiguousarray(lonc)).pin_memory(); h_lat1 = torch.from_numpy(np.ascontiguousarray(latc)).pin_memory()
    cap = int(nx * 1.02) + 1024
    hb = {k: torch.empty(cap, dtype=torch.int32).pin_memory() for k in ("t_in", "i_in", "j_in", "i_out", "j_out")}
    hb.update({k: torch.empty(cap, dtype=torch.float64).pin_memory() for k in (("area", "di", "dj") if order == 2 else ("area",))})
    nx1 = np.full(6, n, np.int32)

    e2e_windows = my_windows if world > 1 else [(bounds[0], bounds[1])]
    h2d_box = [0]

    def e2e_step():
        # what a caller with host grids does every time: the --nlon/--nlat destination is built on the device (nothing to upload),
        # the rank uploads and precomputes only the source rows of its own windows, the result comes back to pinned host memory
        plan.set_dst_latlon(nlon, nlat)
        h2d_box[0] = plan.set_src_sharded(nx1, nx1, h_lon1.numpy().reshape(-1), h_lat1.numpy().reshape(-1), e2e_windows)
        # generate in pieces; each piece is downloaded on a second stream while the next is computed
        return plan.generate_to_host(opcode, hb, nchunks=args.e2e_chunks)

    e2e_steps = max(3, min(args.steps, 10))
    e2e_step()
    barrier()
    with torch.cuda.stream(ext):
        e0.record()
    for _ in range(e2e_steps):
        k = e2e_step()
    with torch.cuda.stream(ext):
        e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d = h2d_box[0]                                   # this rank's share of the source grid (the destination is built on the device)
    h2d_t = torch.tensor([h2d, 0], dtype=torch.int64, device=dev)
    d2h = k * (5 * 4 + (3 if order == 2 else 1) * 8)
    h2d_t[1] = d2h
    if world > 1:
        dist.all_reduce(h2d_t, op=dist.ReduceOp.SUM)
    h2d, d2h = int(h2d_t[0].item()), int(h2d_t[1].item())          # whole job

    apply = None
    if not args.no_apply:
        apply = apply_leg(pkg, torch, dist, rank, world, local, min(args.steps, 10), args.warmup)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    gc = None
    if world == 1 and not args.no_gc:
        gc = gc_leg(pkg, torch, local, min(args.steps, 10), args.warmup, not args.no_cpu_baseline)

    # ---- roofline of the dominant kernel (clip): FP64 pipe, measured DFMA peak
    fp64_peak = pkg.fp64_peak_tflops(local)
    clip_s = float(clip_ms.item()) * 1e-3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    flops = OPS_PER_XCELL[order] * nx_total / max(world, 1)          # per launch (per rank)
    achieved_tf = flops / clip_s * 1e-12 if clip_s > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "clip_traffic.json"))).get(name)
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp64_peak if fp64_peak else None, "traffic": traffic,
                "kernel": f"clip_sh_kernel<{order}> + clip_mom_kernel<{order}> (the clip phase: Sutherland-Hodgman | moments)", "kernel_ms": clip_s * 1e3, "kernel_share_of_step": clip_s * 1e3 / ms,
                "kernel_ms_samples": int(ngen),
                "peak_source": "builder-measured: register-resident DFMA microbenchmark run live by bench.py (csrc/peak_probe.cu); MEASURED_PEAKS.json "
                               "has no FP64 entry and the profiling guide states no FP64 fallback.  The path is compiled -fmad=false (bit-exact "
                               "predicates), so apart from the explicit fma() of the sin/cos routines no DFMA is issued: the reachable ceiling is about half this peak",
                "algorithmic_flops_per_xcell": OPS_PER_XCELL[order],
                "hbm": {"achieved_gbs": (npairs_total / world) * (CLIP_BYTES_PER_PAIR[order] + 0) / clip_s * 1e-9 if clip_s > 0 else None,
                        "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
    phases = {kk: vv / max(ngen, 1) for kk, vv in phase_sum.items()}

    cpu = None
    parity_checked = 0
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        pool = make_pool(name, cores)
        x, tsec, kind, bands = cpu_sample(name, 2, cores, pool, keep=True)
        pool.close()
        cpu = {"value": x / tsec, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{cores} destination row bands x 2 rows of {nlat} (evenly spaced), all 6 source tiles, one process per band: {x} xcells in {tsec:.1f} s"}
        # the CPU sample is also the checker: the end-to-end host result restricted to the sampled rows must BE the CPU lists
        gpu_host = {kk: hb[kk].numpy()[:k] for kk in ("t_in", "i_in", "j_in", "i_out", "j_out", "area")}
        parity_checked = parity_check_bands(gpu_host, bands)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_label(name), "nxgrid": nx_total, "candidate_pairs": npairs_total,
                       "sharding": f"{world * WPR} source-cell windows of equal candidate-pair count dealt round-robin to {world} ranks" if world > 1 else "single window",
                       "step": ("xgb_plan_generate_async per step (+ device-side window counts and NCCL all-gather for N > 1), one "
                                "xgb_plan_generate_finish after the timed region") if use_async else "blocking xgb_plan_generate per step",
                       "l2": "inputs larger than L2 (cell tables ~1.4 GB per rank are re-read every step)"},
            "clocks": clocks, "gpu_launches": launches_total,
            "e2e": {"value": nx_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "api": f"xgb_plan_set_dst_latlon + xgb_plan_set_src_sharded (pinned host source grid, each rank its own rows) + "
                           f"xgb_plan_generate_to_host in {args.e2e_chunks} pieces (pinned host result); bytes are whole-job sums over the ranks"},
            "roofline": roofline, "phase_ms": phases, "per_rank": per_rank, "cpu_baseline": cpu, "apply": apply,
            "parity_checked_xcells": parity_checked, "multi_gpu_parity": multi_parity, "balance": balance, "great_circle": gc}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c768", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--async-steps", action="store_true",
                    help="xgb_plan_generate_async per step + one generate_finish after the timed region instead of the blocking call")
    ap.add_argument("--sync-steps", action="store_true", help="blocking xgb_plan_generate per step also for N > 1 (default there: asynchronous steps)")
    ap.add_argument("--balance-rounds", type=int, default=3,
                    help="N > 1: rounds of measured cost balancing of the source windows before the timed region (0: equal pair counts)")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="pieces of the end-to-end generate (download overlapped with compute)")
    ap.add_argument("--no-apply", action="store_true", help="skip the apply-GB/s leg (configs[1])")
    ap.add_argument("--no-gc", action="store_true", help="skip the great-circle leg (configs[2])")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (sizing runs: tens of GB of pinned memory)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
