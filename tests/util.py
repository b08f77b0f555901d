"""Shared helpers for the parity tests: golden fixtures, oracle parameters, comparison with stated tolerances."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AABB = ((-1.5, -1.5, -1.0), (1.5, 1.5, 1.0))
CFG = {"aabb_min": AABB[0], "aabb_max": AABB[1], "rmnearplane": 0}
H = W = 400
FOCAL = 400.0

# north_star tolerance for floating point outputs: 1e-4 relative (fp32 parity mode).  atol covers values near 0.
RTOL, ATOL = 1e-4, 2e-6


def golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}


def small_params():
    """(P, Pcrf): reference state_dict of the small golden model (tests/golden/params_small.npz)."""
    d = golden("params_small")
    P = {k: v for k, v in d.items() if not k.startswith("crf.")}
    Pc = {k[4:]: v for k, v in d.items() if k.startswith("crf.")}
    return P, Pc


def assert_close(a, b, name, rtol=RTOL, atol=ATOL):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (name, tuple(a.shape), tuple(b.shape))
    if a.numel() == 0:
        return
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bool(bad.any()), (f"{name}: {int(bad.sum())}/{a.numel()} out of tolerance, max|diff|={err.max().item():.3e} "
                                 f"max rel={(err / b.abs().clamp_min(1e-12)).max().item():.3e}")


def synthetic_rays(N, seed=0, n_imgs=30):
    """Same generator as oracle/reference_harness.synthetic_rays."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(N, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 1.0])
    d = torch.cat([torch.randn(N, 2, generator=g) * 0.3, -torch.ones(N, 1)], -1)
    rays = torch.stack([o, d], -1)
    images_idx = torch.randint(0, n_imgs, (N, 1), generator=g)
    return rays, images_idx


def random_params(seed, coarse_grid=(18, 18, 12), fine_grid=(36, 36, 24), device="cpu", scale=0.1):
    """Random parameter dict with the reference's names / shapes for property tests."""
    g = torch.Generator().manual_seed(seed)
    P = {}

    def lin(name, out_c, in_c):
        bound = 1.0 / (in_c ** 0.5)
        P[name] = (torch.rand(out_c, in_c, generator=g) * 2 - 1) * bound

    for pre, (gx, gy, gz), hid, geo in (("mlp_coarse.", coarse_grid, 64, 15), ("mlp_fine.", fine_grid, 256, 128)):
        gs = (gx, gy, gz)
        for i, (m, v) in enumerate((((0, 1), 2), ((0, 2), 1), ((1, 2), 0))):
            c = (64, 16, 16)[i]
            P[pre + f"app_plane.{i}"] = scale * torch.randn(1, c, gs[m[1]], gs[m[0]], generator=g)
            P[pre + f"app_line.{i}"] = scale * torch.randn(1, c, gs[v], 1, generator=g)
        lin(pre + "basis_mat.weight", 32, 96)
        in0 = (32 if pre == "mlp_coarse." else 64) + 63
        lin(pre + "sigma_net.0.weight", hid, in0)
        lin(pre + "sigma_net.1.weight", 1 + geo, hid)
        lin(pre + "color_net.0.weight", hid, geo + 27)
        lin(pre + "color_net.1.weight", hid, hid)
        lin(pre + "color_net.2.weight", 3, hid)
    return {k: v.to(device) for k, v in P.items()}
