"""Multi-GPU plumbing of the exchange-grid path: one process per GPU, torch.distributed (NCCL on the GPUs, gloo in the
CPU tests).  The path shards by contiguous windows of source cells (XgridPlan.partition); a window is a contiguous
range of the reference's emission order, so the rank-order concatenation of the per-rank lists IS the serial list
(the reference's fregrid_parallel gathers in rank order too, conserve_interp.c:404-437, but over destination bands).
The only exchanges are
  * exchange_offsets : all-gather of the per-rank exchange-cell counts -> global offset of each rank's slice;
  * allgather_xgrid  : (apply path) all-gather of the slices themselves, so that every rank holds the whole list and
                       field-levels can be dealt to ranks with no further communication (shard_field_levels).
Nothing here computes geometry; the arrays are whatever XgridPlan.result_device()/result_host() returned."""
import torch
import torch.distributed as dist

INT_KEYS = ("t_in", "i_in", "j_in", "i_out", "j_out")
F64_KEYS = ("area", "di", "dj")


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def exchange_offsets(n_local, device="cpu", group=None, counts_buf=None):
    """-> (offset of this rank's slice in the global list, total, per-rank counts as a list)"""
    rank, world = _world(group)
    if world == 1:
        return 0, int(n_local), [int(n_local)]
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    counts = counts_buf if counts_buf is not None else torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    c = [int(v) for v in counts.tolist()]
    return sum(c[:rank]), sum(c), c


def allgather_xgrid(x_local, group=None):
    """x_local: dict of 1-D tensors (INT_KEYS int32, F64_KEYS float64; di/dj optional) holding this rank's slice.
    Returns the same dict for the whole list, slices concatenated in rank order."""
    rank, world = _world(group)
    keys = [k for k in INT_KEYS + F64_KEYS if x_local.get(k) is not None]
    if world == 1:
        return {k: x_local[k] for k in keys}
    dev = x_local["area"].device
    n = int(x_local["area"].shape[0])
    _, total, counts = exchange_offsets(n, dev, group)
    nmax = max(counts)
    out = {}
    for k in keys:
        v = x_local[k]
        pad = torch.zeros(nmax, dtype=v.dtype, device=dev)
        pad[:n] = v
        buf = torch.empty(world * nmax, dtype=v.dtype, device=dev)
        dist.all_gather_into_tensor(buf, pad, group=group)
        out[k] = torch.cat([buf[r * nmax:r * nmax + counts[r]] for r in range(world)])
    return out


def shard_field_levels(nfields, group=None):
    """field-levels of a batch handled by this rank (round-robin: levels of one time step spread over all GPUs)"""
    rank, world = _world(group)
    return list(range(rank, nfields, world))


def window_bounds_from_counts(pair_counts, nparts, shares=None):
    """Host mirror of the device partition (csrc/xgrid_kernels.cu partition_kernel): bounds[k] = first source cell whose
    exclusive candidate-pair offset reaches k * (total // nparts) — or, with shares (xgb_plan_partition_shares), the
    cumulative share of the total.  pair_counts: 1-D integer tensor/array."""
    c = torch.as_tensor(pair_counts, dtype=torch.int64)
    off = torch.cumsum(c, 0) - c
    total = int(c.sum())
    b = [0]
    run, ssum, prev = 0.0, (float(sum(shares)) if shares is not None else 0.0), 0
    for k in range(1, nparts):
        if shares is None:
            target = (total // nparts) * k
        else:
            run += float(shares[k - 1])
            target = max(prev, min(total, int(float(total) * (run / ssum))))
            prev = target
        b.append(int(torch.searchsorted(off, torch.tensor(target), right=False)))
    b.append(int(c.numel()))
    return b


def rebalance_shares(shares, rank_ms, world, damping=1.0, floor=0.25):
    """Cost-balanced sharding.  Windows are dealt round-robin (window w belongs to rank w % world) and sized by
    candidate-pair count, but a pair is not a fixed cost: the ranks that own a pole spend longer per pair (one warp
    enumerates a pole cap's rows, the pole cell's thousands of exchange cells are summed sequentially to stay
    bit-identical).  Given the device time every rank measured for its share, scale each rank's windows by
    mean(t) / t_rank (damped, bounded below) and renormalise: the next partition gives a slow rank fewer pairs.
    shares: one positive number per window; rank_ms: one time per rank.  Pure arithmetic, identical on every rank."""
    t = [float(v) for v in rank_ms]
    if len(t) != world or min(t) <= 0.0 or len(shares) % world:
        return list(shares)
    mean = sum(t) / world
    out = []
    for w, s in enumerate(shares):
        f = (mean / t[w % world]) ** damping
        out.append(max(float(s) * f, floor * float(s)))
    tot = sum(out)
    return [v / tot for v in out]
