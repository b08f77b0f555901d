"""Ray-batch data parallelism: one process per GPU, the primary-ray batch (with its E exposures and
all samples) is split contiguously across ranks, parameters are replicated.  The forward render needs NO data-path
collective; `gather_rows` only reassembles per-ray outputs when a caller wants the full image on every rank."""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous, balanced partition of n rays: rank r gets [lo, hi); the first n % world ranks get one extra ray."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(t, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def gather_rows(local, n_total, group=None):
    """Inverse of shard_rows: all ranks receive the [n_total, ...] tensor (ragged shards padded to the largest one)."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)


def render_sharded(render_fn, rays, images_idx, group=None):
    """Render this rank's shard of the ray batch with `render_fn(rays, images_idx) -> [n_local, C]` and gather."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(rays.shape[0], rank, world)
    local = render_fn(rays[lo:hi], images_idx[lo:hi])
    return gather_rows(local, rays.shape[0], group)


def max_over_ranks(value, device, group=None):
    """Timing rule for multi-GPU numbers: the slowest rank's device time."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
