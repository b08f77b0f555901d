"""CPU, world_size 2, gloo: the host-side sharding logic of the multi-GPU path (no CUDA kernels involved)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import GOLDEN  # noqa: F401


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from evdeblurnerf_b200.parallel import gather_rows, max_over_ranks, render_sharded, shard_bounds
    rays = torch.arange(n * 6, dtype=torch.float32).reshape(n, 3, 2)
    idx = torch.arange(n).reshape(n, 1)
    fake_render = lambda r, i: torch.cat([r.reshape(r.shape[0], -1)[:, :2], i.float()], -1) * 2.0   # per-ray function
    full = render_sharded(fake_render, rays, idx)
    assert torch.equal(full, fake_render(rays, idx)), "sharded render + gather must equal the single-process render"
    lo, hi = shard_bounds(n, rank, world)
    assert torch.equal(gather_rows(rays[lo:hi], n), rays)
    assert max_over_ranks(float(rank + 1), "cpu") == float(world)
    dist.destroy_process_group()


def test_shard_bounds_cover_exactly_once():
    from evdeblurnerf_b200.parallel import shard_bounds
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_sharded_render_world2_gloo():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 4097), nprocs=2, join=True)
