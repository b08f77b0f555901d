"""GPU parity of the backward pass: edn_render_field_bwd (through the C ABI) against torch autograd on the oracle.

Tolerance: every gradient tensor within 1e-4 of its own max magnitude plus 1e-4 relative (fp32 parity mode; the backward
reductions over samples run in a different order than autograd's, and the scatter-adds are unordered atomics)."""
import pytest
import torch

import evdeblur_oracle as oc
from util import AABB, CFG, FOCAL, H, W, assert_close, oracle_fine_at, random_params, synthetic_rays

pytestmark = pytest.mark.gpu


def grad_close(a, b, name, tol=1e-4):
    b = torch.as_tensor(b)
    scale = float(b.abs().max())
    assert scale > 0, f"{name}: oracle gradient is identically zero (test is vacuous)"
    assert_close(a, b, name, rtol=tol, atol=tol * scale)


def setup(seed, R, bias=False):
    P = random_params(seed, scale=0.3)
    if bias:
        g = torch.Generator().manual_seed(seed + 1)
        for pre, hid in (("mlp_coarse.", 64), ("mlp_fine.", 256)):
            P[pre + "color_net.0.bias"] = 0.1 * torch.randn(hid, generator=g)
            P[pre + "color_net.1.bias"] = 0.1 * torch.randn(hid, generator=g)
            P[pre + "color_net.2.bias"] = 0.1 * torch.randn(3, generator=g)
    rays, _ = synthetic_rays(R, seed=seed)
    rb = oc.build_ray_batch(H, W, FOCAL, rays)
    return P, rb


def cotangents(seed, R):
    g = torch.Generator().manual_seed(1000 + seed)
    return {k: torch.randn(R, *shp, generator=g) for k, shp in
            (("rgb_map", (3,)), ("depth_map", ()), ("acc_map", ()), ("rgb0", (3,)), ("depth0", ()), ("acc0", ()))}


def oracle_grads(P, rb, Nc, z_all, cot, noise0=None, noise1=None):
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    rbg = rb.clone().requires_grad_(True)
    c = oc.render_rays(Pg, CFG, rbg, Nc, 0, rand={"noise0": noise0})
    if z_all is None:
        loss = (c["rgb_map"] * cot["rgb_map"]).sum() + (c["depth_map"] * cot["depth_map"]).sum() + (c["acc_map"] * cot["acc_map"]).sum()
    else:
        f = oracle_fine_at(Pg, rbg, z_all, noise1)
        loss = (c["rgb_map"] * cot["rgb0"]).sum() + (c["depth_map"] * cot["depth0"]).sum() + (c["acc_map"] * cot["acc0"]).sum()
        loss = loss + (f["rgb_map"] * cot["rgb_map"]).sum() + (f["depth_map"] * cot["depth_map"]).sum() + (f["acc_map"] * cot["acc_map"]).sum()
    names = [k for k in Pg if z_all is not None or k.startswith("mlp_coarse.")]
    gs = torch.autograd.grad(loss, [Pg[k] for k in names] + [rbg], allow_unused=True)
    out = {k: g for k, g in zip(names, gs[:-1])}
    return out, gs[-1]


@pytest.mark.parametrize("bias", [False, True])
def test_coarse_field_backward_matches_autograd(bias):
    from evdeblurnerf_b200 import RenderEngine
    from evdeblurnerf_b200.backward import render_rays_backward
    R, Nc = 80, 48
    P, rb = setup(3, R, bias)
    cot = cotangents(3, R)
    noise0 = 0.5 * torch.randn(R, Nc - 1, generator=torch.Generator().manual_seed(5))
    ref, ref_rb = oracle_grads(P, rb, Nc, None, cot, noise0=noise0)
    eng = RenderEngine({k: v.cuda() for k, v in P.items() if k.startswith("mlp_coarse.")}, *AABB, precision="fp32")
    out = eng.render_rays(rb.cuda(), Nc, retraw=True, rand={"noise0": noise0.cuda()})
    saved = {"ray_batch": rb.cuda(), "z_vals0": out["z_vals"], "noise0": noise0.cuda()}
    grads, d_rb = render_rays_backward(eng, saved, {k: cot[k].cuda() for k in ("rgb_map", "depth_map", "acc_map")}, chunk_rays=32)
    got = grads.finish()
    for k, g in ref.items():
        grad_close(got[k], g, k)
    grad_close(d_rb[:, :6], ref_rb[:, :6], "d ray_batch[o, d]")
    grad_close(d_rb[:, 8:], ref_rb[:, 8:], "d ray_batch[viewdirs]")
    assert float(d_rb[:, 6:8].abs().max()) == 0.0


@pytest.mark.parametrize("merge", [False, True])
def test_c2f_backward_matches_autograd(merge):
    """merge = True: the merged order of the forward is handed to the backward, which then scatters every coarse position into the
    coarse grid once (edn_field_bwd_merge) -- same gradients."""
    from evdeblurnerf_b200 import RenderEngine
    from evdeblurnerf_b200.backward import render_rays_backward
    R, Nc, Ni = 64, 32, 32
    P, rb = setup(4, R, bias=True)
    cot = cotangents(4, R)
    g = torch.Generator().manual_seed(6)
    noise0, noise1 = 0.5 * torch.randn(R, Nc - 1, generator=g), 0.5 * torch.randn(R, Nc + Ni - 1, generator=g)
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision="fp32")
    rand = {"noise0": noise0.cuda(), "noise1": noise1.cuda()}
    out = eng.render_rays(rb.cuda(), Nc, retraw=True, N_importance=Ni, rand=rand, want_indices=merge)
    ref, ref_rb = oracle_grads(P, rb, Nc, out["z_vals"].cpu(), cot, noise0=noise0, noise1=noise1)
    saved = {"ray_batch": rb.cuda(), "z_vals0": out["z_vals0"], "z_vals": out["z_vals"], "noise0": rand["noise0"], "noise1": rand["noise1"],
             "order": out["order"] if merge else None}
    grads, d_rb = render_rays_backward(eng, saved, {k: v.cuda() for k, v in cot.items()}, chunk_rays=24)
    got = grads.finish()
    for k, gr in ref.items():
        grad_close(got[k], gr, k)
    grad_close(d_rb[:, :6], ref_rb[:, :6], "d ray_batch[o, d]")
    grad_close(d_rb[:, 8:], ref_rb[:, 8:], "d ray_batch[viewdirs]")


def test_backward_accumulates_and_chunking_is_invisible():
    from evdeblurnerf_b200 import RenderEngine
    from evdeblurnerf_b200.backward import RenderGradients, render_rays_backward
    R, Nc, Ni = 50, 32, 32
    P, rb = setup(7, R)
    cot = {k: v.cuda() for k, v in cotangents(7, R).items()}
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision="fp32")
    out = eng.render_rays(rb.cuda(), Nc, retraw=True, N_importance=Ni)
    saved = {"ray_batch": rb.cuda(), "z_vals0": out["z_vals0"], "z_vals": out["z_vals"]}
    g1, d1 = render_rays_backward(eng, saved, cot, chunk_rays=7)
    g2, d2 = render_rays_backward(eng, saved, cot, chunk_rays=4096)
    a, b = g1.finish(), g2.finish()
    for k in a:
        grad_close(a[k], b[k], "chunk " + k, tol=2e-5)
    grad_close(d1, d2, "chunk d_rb", tol=2e-5)
    # the merged coarse-position scatter (edn_field_bwd_merge) is a re-association of the same sums, chunked or not
    out_i = eng.render_rays(rb.cuda(), Nc, retraw=True, N_importance=Ni, want_indices=True)
    assert torch.equal(out_i["z_vals"], out["z_vals"])
    saved_m = dict(saved, order=out_i["order"])
    for ch in (7, 4096):
        g3, d3 = render_rays_backward(eng, saved_m, cot, chunk_rays=ch)
        c3 = g3.finish()
        for k in c3:
            grad_close(c3[k], b[k], f"merged scatter (chunk {ch}) " + k, tol=2e-5)
        grad_close(d3, d2, f"merged scatter (chunk {ch}) d_rb", tol=2e-5)
    acc = RenderGradients(eng)
    render_rays_backward(eng, saved, cot, grads=acc)
    render_rays_backward(eng, saved, cot, grads=acc)
    c = acc.finish()
    for k in ("mlp_fine.sigma_net.0.weight", "mlp_coarse.app_plane.0", "mlp_fine.basis_mat.weight"):
        grad_close(c[k], 2 * b[k], "accumulate " + k, tol=2e-5)


def test_bf16_mode_gradients_are_close_to_the_fp32_oracle():
    """bf16 mode: bf16 VM planes, bf16 activation storage and GEMM operands in the backward (fp32 accumulation, fp32 weight
    gradients).  Not a parity mode: every gradient tensor must point the same way as the fp32 oracle's (cosine > 0.995) with a
    relative L2 error under 10 %."""
    from evdeblurnerf_b200 import RenderEngine
    from evdeblurnerf_b200.backward import render_rays_backward
    R, Nc, Ni = 64, 32, 32
    P, rb = setup(9, R, bias=True)
    cot = cotangents(9, R)
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision="bf16")
    out = eng.render_rays(rb.cuda(), Nc, retraw=True, N_importance=Ni)
    ref, ref_rb = oracle_grads(P, rb, Nc, out["z_vals"].cpu(), cot)
    saved = {"ray_batch": rb.cuda(), "z_vals0": out["z_vals0"], "z_vals": out["z_vals"]}
    grads, d_rb = render_rays_backward(eng, saved, {k: v.cuda() for k, v in cot.items()}, chunk_rays=40)
    got = grads.finish()
    worst = {}
    keep = [0, 1, 2, 3, 4, 5, 8, 9, 10]          # near / far are constants of the render path (no gradient by design)
    for k, g in list(ref.items()) + [("d_ray_batch", ref_rb[:, keep])]:
        a = (d_rb[:, keep] if k == "d_ray_batch" else got[k]).detach().cpu().double().flatten()
        b = g.double().flatten()
        cos = float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))
        rel = float((a - b).norm() / b.norm().clamp_min(1e-30))
        worst[k] = (cos, rel)
    bad = {k: v for k, v in worst.items() if v[0] < 0.995 or v[1] > 0.10}
    assert not bad, bad


def test_backward_error_paths_and_empty_batch():
    """Error behaviour mirrors the reference's exceptions: a too small workspace or a missing gradient buffer raises with
    the library's message; an empty ray batch is a no-op."""
    import ctypes as C
    from evdeblurnerf_b200 import RenderEngine, _lib
    from evdeblurnerf_b200.backward import RenderGradients, _field_struct, render_rays_backward
    P, rb = setup(11, 8)
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision="fp32")
    out = eng.render_rays(rb.cuda(), 32, retraw=True, N_importance=32)
    saved = {"ray_batch": rb.cuda(), "z_vals0": out["z_vals0"], "z_vals": out["z_vals"]}
    # empty batch
    empty = {"ray_batch": rb.cuda()[:0], "z_vals0": out["z_vals0"][:0], "z_vals": out["z_vals"][:0]}
    grads, d_rb = render_rays_backward(eng, empty, {})
    assert d_rb.shape == (0, 11) and all(float(v.abs().max()) == 0.0 for v in grads.finish().values())
    # workspace too small
    lib = _lib.load()
    g = RenderGradients(eng)
    w = _field_struct(eng.params, "mlp_fine.", ["mlp_coarse.", "mlp_fine."])
    gw = _field_struct(g.w, "mlp_fine.", ["mlp_coarse.", "mlp_fine."])
    d_rbuf = torch.zeros(8, 11).cuda()
    ws = torch.empty(1024, dtype=torch.uint8).cuda()
    args = lambda gw_, ws_, n_: (C.byref(eng.coarse.grid), C.byref(eng.fine.grid), C.byref(w), saved["ray_batch"].data_ptr(),
                                 saved["z_vals"].data_ptr(), None, 8, 64, 0, None, None, None, None, None, C.byref(gw_),
                                 C.byref(g.grid_struct["mlp_coarse."]), C.byref(g.grid_struct["mlp_fine."]), d_rbuf.data_ptr(), ws_.data_ptr(), n_,
                                 None, torch.cuda.current_stream().cuda_stream)
    with pytest.raises(RuntimeError, match="workspace too small"):
        _lib.check(lib.edn_render_field_bwd(*args(gw, ws, 1024)), "edn_render_field_bwd")
    gw.sigma0 = None
    big = torch.empty(int(lib.edn_field_bwd_workspace_bytes(2, 256, 128, 8, 64)), dtype=torch.uint8).cuda()
    with pytest.raises(RuntimeError, match="null weight gradient"):
        _lib.check(lib.edn_render_field_bwd(*args(gw, big, big.numel())), "edn_render_field_bwd")
