"""Multi-GPU plumbing: one process per GPU (torchrun), instances statically sharded across ranks, no
collective on the data path, one final gather of (tours, costs) over NCCL/NVLink (gloo on CPU for
the host-logic tests).  The reference has no distributed code; this is the scale-out BASELINE.json
asks for (instances are independent, scripts/test.py:59 processes them one by one)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous static shard [lo, hi) of `total` instances for `rank`; sizes differ by at most 1."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(tours, costs, total=None, group=None):
    """all_gather this rank's (tours [b,n+1] int32, costs [b] fp64) into ([B,n+1], [B]) on every rank, in
    global instance order.  Shards may be uneven (padded to the largest shard for the collective)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tours, costs
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if total is None:
        t = torch.tensor([tours.shape[0]], dtype=torch.int64, device=tours.device)
        dist.all_reduce(t, group=group)
        total = int(t)
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    if tours.shape[0] != sizes[rank]:
        raise ValueError(f'rank {rank} holds {tours.shape[0]} instances, expected {sizes[rank]}')
    m, n1 = max(sizes), tours.shape[1]
    pt = torch.zeros(m, n1, dtype=tours.dtype, device=tours.device)
    pc = torch.zeros(m, dtype=costs.dtype, device=costs.device)
    pt[: tours.shape[0]] = tours
    pc[: costs.shape[0]] = costs
    gt = torch.empty(world * m, n1, dtype=tours.dtype, device=tours.device)
    gc = torch.empty(world * m, dtype=costs.dtype, device=costs.device)
    dist.all_gather_into_tensor(gt, pt, group=group)
    dist.all_gather_into_tensor(gc, pc, group=group)
    if all(s == m for s in sizes):
        return gt, gc
    keep = torch.cat([torch.arange(r * m, r * m + s, device=tours.device) for r, s in enumerate(sizes)])
    return gt[keep], gc[keep]


def solve_sharded(solver, D_all_host, **kw):
    """Every rank solves its shard of D_all_host [B,n,n] (host tensor/array) on its own GPU and returns the
    gathered (tours, costs) for all B instances (device tensors)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    D_all_host = torch.as_tensor(D_all_host)
    lo, hi = shard_range(D_all_host.shape[0], rank, world)
    dev = next(solver.model.parameters()).device
    res = solver.solve(D_all_host[lo:hi].to(dev), **kw)
    return gather_results(res.best_tours, res.best_costs, D_all_host.shape[0])
