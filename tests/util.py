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


ACHIEVED = {}      # test id -> {comparison name: achieved error}: written to gpurun_out/parity_achieved.json at session end (conftest.py)


def assert_close(a, b, name, rtol=RTOL, atol=ATOL):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (name, tuple(a.shape), tuple(b.shape))
    if a.numel() == 0:
        return
    err = (a - b).abs()
    # the ACHIEVED error is recorded beside the bound (max |diff|, max |diff| / max |ref|, and the worst ratio err / allowed)
    test_id = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
    ACHIEVED.setdefault(test_id, {})[name] = {
        "max_abs": float(err.max()), "max_abs_over_ref_max": float(err.max() / b.abs().max().clamp_min(1e-30)),
        "worst_err_over_allowed": float((err / (atol + rtol * b.abs())).max()), "rtol": rtol, "atol": atol}
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


def oracle_fine_at(P, ray_batch, z_all, noise=None, is_train=True, rmnearplane=0):
    """Oracle fine stage (renderer.py:206-217) evaluated at GIVEN merged depths: sample_pdf is ill-conditioned in the
    last bits of weights0 (t = (u - cdf_b) / denom with denom down to 1e-5), so the fine pass is checked tightly at the
    depths the CUDA path itself produced, and the sampler separately (bit-exact on identical weights)."""
    import evdeblur_oracle as oc
    o, d, vd = ray_batch[:, 0:3], ray_batch[:, 3:6], ray_batch[:, -3:]
    pts = o[:, None, :] + d[:, None, :] * z_all[..., None]
    ft = torch.cat([oc.vm_sample(P, "mlp_coarse.", pts, *AABB), oc.vm_sample(P, "mlp_fine.", pts, *AABB)], -1)
    rgb, depth, acc, w, feat = oc.field_forward(P, "mlp_fine.", pts, vd, ft, z_all, d, noise, is_train, rmnearplane,
                                                rgb_act="none")
    return {"rgb_map": rgb, "depth_map": depth, "acc_map": acc, "weights": w, "depth_feature": feat}


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def emulated_bf16_fine(P, ray_batch, z_all, noise=None, is_train=True, rmnearplane=0, lean=False):
    """Torch fp32 reference of the tcgen05 fine pass with its operand roundings made explicit: VM planes, the
    plane (.) line products, every MMA A operand (features, PE, activations) and every MMA weight matrix are rounded to
    bf16; accumulation, biases, the sigma / rgb heads (fp32 dot products in the epilogues), sigmoid and compositing stay fp32.  The view-direction part of color_net.0 is an
    fp32 per-ray bias.  Against this reference the CUDA kernel differs by accumulation order only."""
    import evdeblur_oracle as oc
    import torch.nn.functional as F
    o, d, vd = ray_batch[:, 0:3], ray_batch[:, 3:6], ray_batch[:, -3:]
    R, S = z_all.shape
    pts = o[:, None, :] + d[:, None, :] * z_all[..., None]
    Pb = {k: (_bf(v) if ("app_plane" in k or "app_line" in k) else v) for k, v in P.items()}
    fts, gs = [], []
    for pre in ("mlp_coarse.", "mlp_fine."):
        g = _bf(oc.vm_products(Pb, pre, pts, *AABB))
        gs.append(g)
        fts.append(_bf(F.linear(g, _bf(P[pre + "basis_mat.weight"]))))
    pe = _bf(oc.posenc(pts.reshape(-1, 3), 10))
    pre = "mlp_fine."
    w0 = P[pre + "sigma_net.0.weight"]
    if lean:   # "lean" schedule: basis_mat folded into sigma_net.0 in fp32, the product rounded to bf16 once
        w0f = torch.cat([w0[:, :32] @ P["mlp_coarse.basis_mat.weight"], w0[:, 32:64] @ P["mlp_fine.basis_mat.weight"], w0[:, 64:]], 1)
        h1_f32 = torch.relu(F.linear(torch.cat(gs + [pe], -1), _bf(w0f)))
    else:
        h1_f32 = torch.relu(F.linear(torch.cat(fts + [pe], -1), _bf(w0)))
    h1 = _bf(h1_f32)
    w1 = P[pre + "sigma_net.1.weight"]
    sigma = F.linear(h1_f32, w1[:1])            # sigma head: fp32 dot product in the layer epilogue
    geo = F.linear(h1, _bf(w1[1:]))
    w3 = P[pre + "color_net.0.weight"]
    bias_ray = F.linear(oc.posenc(vd, 4), w3[:, 128:], P.get(pre + "color_net.0.bias"))          # [R,256] fp32
    if lean:   # sigma_net.1 (geo columns) folded into color_net.0
        h3 = F.linear(h1, _bf(w3[:, :128] @ w1[1:])) + bias_ray[:, None, :].expand(R, S, 256).reshape(R * S, 256)
    else:
        h3 = F.linear(_bf(geo), _bf(w3[:, :128])) + bias_ray[:, None, :].expand(R, S, 256).reshape(R * S, 256)
    h3 = _bf(torch.relu(h3))
    h4 = torch.relu(F.linear(h3, _bf(P[pre + "color_net.1.weight"]), P.get(pre + "color_net.1.bias")))
    rgb = torch.sigmoid(F.linear(h4, P[pre + "color_net.2.weight"], P.get(pre + "color_net.2.bias")))   # fp32 rgb head
    raw = torch.cat([sigma, rgb], -1).reshape(R, S, 4)
    rgb_map, _, acc, w, depth = oc.raw2outputs(raw, z_all, d, noise, is_train, rmnearplane)
    return {"rgb_map": rgb_map, "depth_map": depth, "acc_map": acc, "weights": w, "depth_feature": geo.reshape(R, S, -1),
            "sigma": sigma.reshape(R, S)}


def emulated_bf16_coarse(P, ray_batch, n_samples, t_rand=None, noise=None, is_train=True, rmnearplane=0, lean=True):
    """Torch fp32 reference of the tcgen05 coarse pass with its bf16 operand roundings made explicit (same
    conventions as emulated_bf16_fine; sigma comes out of the sigma_net.1 MMA, the rgb head is an fp32 dot product)."""
    import evdeblur_oracle as oc
    import torch.nn.functional as F
    o, d, vd = ray_batch[:, 0:3], ray_batch[:, 3:6], ray_batch[:, -3:]
    z = oc.place_samples(ray_batch[:, 6:7], ray_batch[:, 7:8], n_samples, False, t_rand)
    R, S = z.shape
    pts = o[:, None, :] + d[:, None, :] * z[..., None]
    pre = "mlp_coarse."
    Pb = {k: (_bf(v) if ("app_plane" in k or "app_line" in k) else v) for k, v in P.items()}
    g = _bf(oc.vm_products(Pb, pre, pts, *AABB))
    pe = _bf(oc.posenc(pts.reshape(-1, 3), 10))
    w0, w1, w3 = P[pre + "sigma_net.0.weight"], P[pre + "sigma_net.1.weight"], P[pre + "color_net.0.weight"]
    bias_ray = F.linear(oc.posenc(vd, 4), w3[:, 15:], P.get(pre + "color_net.0.bias"))
    if lean:   # default schedule: basis_mat folded into sigma_net.0, sigma_net.1's geo rows folded into color_net.0, fp32 sigma head
        w0f = torch.cat([w0[:, :32] @ P[pre + "basis_mat.weight"], w0[:, 32:]], 1)
        h1_f32 = torch.relu(F.linear(torch.cat([g, pe], -1), _bf(w0f)))
        h1 = _bf(h1_f32)
        sigma = F.linear(h1_f32, w1[:1])
        geo = F.linear(h1, w1[1:])
        c0 = F.linear(h1, _bf(w3[:, :15] @ w1[1:]))
    else:
        ft = _bf(F.linear(g, _bf(P[pre + "basis_mat.weight"])))
        h1 = _bf(torch.relu(F.linear(torch.cat([ft, pe], -1), _bf(w0))))
        o16 = F.linear(h1, _bf(w1))
        sigma, geo = o16[:, :1], o16[:, 1:]
        c0 = F.linear(_bf(geo), _bf(w3[:, :15]))
    c0 = c0 + bias_ray[:, None, :].expand(R, S, 64).reshape(R * S, 64)
    c0 = _bf(torch.relu(c0))
    c1 = torch.relu(F.linear(c0, _bf(P[pre + "color_net.1.weight"]), P.get(pre + "color_net.1.bias")))
    rgb = torch.sigmoid(F.linear(c1, P[pre + "color_net.2.weight"], P.get(pre + "color_net.2.bias")))
    raw = torch.cat([sigma, rgb], -1).reshape(R, S, 4)
    rgb_map, _, acc, w, depth = oc.raw2outputs(raw, z, d, noise, is_train, rmnearplane, rgb_act="relu")
    return {"rgb_map": rgb_map, "depth_map": depth, "acc_map": acc, "weights": w, "z_vals": z, "feature": geo.reshape(R, S, -1)}
