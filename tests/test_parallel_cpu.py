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


def _grad_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from evdeblurnerf_b200.trainer import FlatParams
    g = torch.Generator().manual_seed(0)
    tensors = {"mlp_fine.color_net.1.weight": torch.randn(5, 7, generator=g), "mlp_fine.sigma_net.0.weight": torch.randn(3, 3, generator=g),
               "mlp_coarse.app_plane.0": torch.randn(1, 2, 3, 5, generator=g), "kernelsnet.r_linear.bias": torch.randn(6, generator=g)}
    flat = FlatParams(tensors, "cpu")
    for k, v in flat.views.items():
        assert torch.equal(v.detach(), tensors[k]) and v.grad.data_ptr() >= flat.grad.data_ptr()
        assert v.data_ptr() % 16 == 0 or (v.data_ptr() - flat.param.data_ptr()) % 16 == 0
    # rank-dependent "loss": autograd accumulates straight into the flat gradient buffer through the views
    loss = sum(((rank + 1) * (i + 1)) * (v ** 2).sum() for i, v in enumerate(flat.views[k] for k in sorted(flat.views)))
    loss.backward()
    flat.all_reduce_mean()
    mean_scale = sum(r + 1 for r in range(world)) / world
    for i, k in enumerate(sorted(flat.views)):
        assert torch.allclose(flat.views[k].grad, 2 * mean_scale * (i + 1) * tensors[k], rtol=1e-6, atol=1e-6), k
    # weight-decayed tensors (color_net.N.weight, run_nerf.py:246) form the leading segment
    assert flat.order[0] == "mlp_fine.color_net.1.weight" and flat.split == 36
    dist.destroy_process_group()


def test_flat_gradient_all_reduce_world2_gloo():
    port = _free_port()
    mp.spawn(_grad_worker, args=(2, port), nprocs=2, join=True)


def test_lr_schedule_matches_reference_formula():
    from evdeblurnerf_b200.trainer import lr_at
    # run_nerf.py:604-613 with lrate 5e-4, lrate_decay 250, warm-up 2000 its from factor 0.1
    assert abs(lr_at(0, 5e-4, 250, 2000, 0.1) - 5e-5) < 1e-12
    assert abs(lr_at(1000, 5e-4, 250, 2000, 0.1) - 5e-4 * 0.55) < 1e-12
    assert abs(lr_at(250000, 5e-4, 250, 2000, 0.1) - 5e-5) < 1e-12
    assert abs(lr_at(0, 5e-4, 250) - 5e-4) < 1e-12
