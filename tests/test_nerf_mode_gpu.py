"""GPU parity of mode = nerf ("run_network") against the golden vectors of the unmodified reference
and the oracle.  fp32, 1e-4 relative."""
import pytest
import torch

import evdeblur_oracle as oc
from util import AABB, assert_close, golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case5():
    g = golden("case5_nerf24")
    Pn = {k[2:]: v for k, v in g.items() if k.startswith("P.")}
    return g, Pn


def test_nerf_mlpforward_and_raw2outputs_golden(case5):
    from evdeblurnerf_b200.nerf_mode import NeRF
    g, Pn = case5
    net = NeRF({k: v.cuda() for k, v in Pn.items()}, "mlp_fine.")
    rb, z = g["ray_batch"].cuda(), g["z_vals"].cuda()
    raw, feat = net.mlpforward_at(rb, z, want_feature=True)
    assert_close(raw, g["raw"], "raw", rtol=1e-4, atol=2e-5)
    assert_close(feat.double().sum(-1).float(), g["feature_sum"], "feature (before_linear) row sums", rtol=1e-4, atol=1e-3)
    rgb, depth, acc, w = net.raw2outputs(raw, z, rb)
    assert_close(rgb, g["rgb_map"], "rgb_map", rtol=1e-4, atol=2e-5)
    assert_close(depth, g["depth_map"], "depth_map", rtol=1e-4, atol=2e-5)
    assert_close(acc, g["acc_map"], "acc_map", rtol=1e-4, atol=2e-5)
    assert_close(w, g["weights"], "weights", rtol=1e-4, atol=2e-5)
    o = oc.nerf_raw2outputs(g["raw"], g["z_vals"], g["ray_batch"][:, 3:6], white_bkgd=True)
    rgb_w, _, _, _ = net.raw2outputs(g["raw"].cuda(), z, rb, white_bkgd=True)
    assert_close(rgb_w, o[0], "white background", rtol=1e-4, atol=2e-5)


def test_nerf_render_rays_c2f_sampling(case5):
    from evdeblurnerf_b200 import RenderEngine
    from evdeblurnerf_b200.nerf_mode import NeRF, render_rays_nerf
    from util import small_params
    g, Pn = case5
    Pg = {k: v.cuda() for k, v in Pn.items()}
    net = NeRF(Pg, "mlp_fine.")
    P, _ = small_params()
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision="fp32")     # sampler host of the nerf-mode render
    rb = g["ray_batch"]
    out = render_rays_nerf(eng, net, net, rb.cuda(), 64, retraw=True, N_importance=64)
    # oracle: same field for the coarse and the fine pass
    z0 = oc.place_samples(rb[:, 6:7], rb[:, 7:8], 64)
    assert torch.equal(out["z_vals0"].cpu(), z0)
    o, d, vd = rb[:, :3], rb[:, 3:6], rb[:, -3:]
    raw0, _ = oc.nerf_mlpforward(Pn, "mlp_fine.", o[:, None] + d[:, None] * z0[..., None], vd)
    r0 = oc.nerf_raw2outputs(raw0, z0, d)
    assert_close(out["rgb0"], r0[0], "rgb0", rtol=1e-4, atol=2e-5)
    assert_close(out["weights0"], r0[3], "weights0", rtol=1e-4, atol=2e-5)
    zf = out["z_vals"].cpu()
    raw1, _ = oc.nerf_mlpforward(Pn, "mlp_fine.", o[:, None] + d[:, None] * zf[..., None], vd)
    r1 = oc.nerf_raw2outputs(raw1, zf, d)
    assert_close(out["rgb_map"], r1[0], "rgb_map at the CUDA depths", rtol=1e-4, atol=2e-5)
    assert_close(out["weights"], r1[3], "weights", rtol=1e-4, atol=2e-5)
    assert_close(out["rgb_map"], g["rgb_map"], "golden rgb_map (end to end)", rtol=1e-4, atol=5e-4)


def test_nerf_ragged(case5):
    from evdeblurnerf_b200.nerf_mode import NeRF
    g, Pn = case5
    net = NeRF({k: v.cuda() for k, v in Pn.items()}, "mlp_fine.", extract_feature="after_linear")
    rb = g["ray_batch"][:5]
    z = torch.sort(torch.rand(5, 37, generator=torch.Generator().manual_seed(1)), -1)[0]
    raw, feat = net.mlpforward_at(rb.cuda(), z.cuda(), want_feature=True)
    o, d, vd = rb[:, :3], rb[:, 3:6], rb[:, -3:]
    raw_o, feat_o = oc.nerf_mlpforward(Pn, "mlp_fine.", o[:, None] + d[:, None] * z[..., None], vd, before_linear=False)
    assert_close(raw, raw_o, "raw", rtol=1e-4, atol=2e-5)
    assert_close(feat, feat_o, "feature (after_linear)", rtol=1e-4, atol=2e-5)


def test_nerfall_facade_nerf_mode(case5):
    from evdeblurnerf_b200 import NeRFAll
    g, Pn = case5
    Pg = {k: v.cuda() for k, v in Pn.items()}
    Pg.update({k.replace("mlp_fine.", "mlp_coarse."): v for k, v in Pg.items()})
    nerf = NeRFAll(Pg, *AABB).train()
    assert nerf.mode == "nerf"
    out = nerf.render_rays(g["ray_batch"].cuda(), 64, retraw=True, N_importance=64)
    assert_close(out["rgb_map"], g["rgb_map"], "facade rgb_map", rtol=1e-4, atol=5e-4)


def _grad_close(a, b, name, tol=2e-4):
    scale = float(torch.as_tensor(b).abs().max())
    assert scale > 0, f"{name}: oracle gradient is identically zero"
    assert_close(a, b, name, rtol=tol, atol=tol * scale)


@pytest.mark.parametrize("before_linear,white_bkgd", [(True, False), (False, True)])
def test_nerf_field_backward_matches_autograd(case5, before_linear, white_bkgd):
    """edn_nerf_field_bwd against torch autograd on the oracle: all 24 tensors of the field, d ray_batch; upstream gradients on
    rgb / depth / acc and on the extracted feature (both extract_feature modes), white background on / off; 2e-4 of max."""
    from evdeblurnerf_b200 import NerfRenderEngine
    from evdeblurnerf_b200.nerf_mode import NeRF
    g, Pn = case5
    gen = torch.Generator().manual_seed(12)
    Pn = dict(Pn)
    Pn["mlp_fine.rgb_linear.bias"] = 0.1 * torch.randn(3, generator=gen)
    R, S = 20, 45
    rb = g["ray_batch"][:R]
    z = torch.sort(torch.rand(R, S, generator=gen), -1)[0]
    noise = 0.3 * torch.randn(R, S - 1, generator=gen)
    cot = {"rgb": torch.randn(R, 3, generator=gen), "depth": torch.randn(R, generator=gen), "acc": torch.randn(R, generator=gen),
           "feat": 0.05 * torch.randn(R, S, 256, generator=gen)}
    Po = {k: v.clone().requires_grad_(True) for k, v in Pn.items()}
    rbo = rb.clone().requires_grad_(True)
    o, d, vd = rbo[:, :3], rbo[:, 3:6], rbo[:, -3:]
    raw, feat = oc.nerf_mlpforward(Po, "mlp_fine.", o[:, None] + d[:, None] * z[..., None], vd, before_linear=before_linear)
    rgb, _, acc, _, depth = oc.nerf_raw2outputs(raw, z, d, noise, white_bkgd=white_bkgd)
    ((rgb * cot["rgb"]).sum() + (depth * cot["depth"]).sum() + (acc * cot["acc"]).sum() + (feat * cot["feat"]).sum()).backward()

    eng = NerfRenderEngine({k.replace("mlp_fine.", "mlp_coarse."): v.cuda() for k, v in Pn.items()}, use_awp=before_linear)
    net = NeRF({k: v.cuda() for k, v in Pn.items()}, "mlp_fine.", "before_linear" if before_linear else "after_linear")
    d_rb = torch.zeros(R, 11).cuda()
    grads = net.backward(eng, rb.cuda(), z.cuda(), noise.cuda(), cot["rgb"].cuda(), cot["depth"].cuda(), cot["acc"].cuda(), d_rb,
                         white_bkgd=white_bkgd, d_feat=cot["feat"].cuda(), chunk_rays=7)
    for k in Po:
        _grad_close(grads[k], Po[k].grad, k)
    _grad_close(d_rb[:, :6], rbo.grad[:, :6], "d ray_batch[o, d]")
    _grad_close(d_rb[:, 8:], rbo.grad[:, 8:], "d ray_batch[viewdirs]")


def test_nerfall_nerf_mode_loss_backward(case5):
    """loss.backward() through the facade in mode = nerf (coarse + fine fields): gradients reach every tensor of both fields."""
    from evdeblurnerf_b200 import NeRFAll, img2mse
    g, Pn = case5
    Pg = {k: v.clone().cuda().requires_grad_(True) for k, v in Pn.items()}
    Pg.update({k.replace("mlp_fine.", "mlp_coarse."): v.detach().clone().requires_grad_(True) for k, v in list(Pg.items())})
    nerf = NeRFAll(Pg, *AABB).train()
    R = 16
    rays = torch.stack([g["ray_batch"][:R, :3], g["ray_batch"][:R, 3:6]], -1).cuda()
    rgb, rgb0, _, _ = nerf(400, 400, [[400.0, 0, 200.0], [0, 400.0, 200.0], [0, 0, 1.0]], rays=rays, rays_info=None, force_naive=True,
                           ndc=False, N_samples=32, N_importance=32, perturb=0., raw_noise_std=0.)
    target = torch.rand(R, 3, generator=torch.Generator().manual_seed(2)).cuda()
    (img2mse(rgb, target) + img2mse(rgb0, target)).backward()
    for k, v in Pg.items():
        assert v.grad is not None and float(v.grad.abs().max()) > 0, k


def nerf_awp_model(requires_grad):
    """mode = nerf + RBK + AWP on 256-channel features: case 5's MLP for both passes, kernel / AWP nets of case 10."""
    from evdeblurnerf_b200 import NeRFAll
    g5, g10 = golden("case5_nerf24"), golden("case10_nerf_awp")
    P = {}
    for k, v in g5.items():
        if k.startswith("P.mlp_fine."):
            P[k[2:]] = v
            P["mlp_coarse." + k[len("P.mlp_fine."):]] = v.clone()
    P.update({k[2:]: v for k, v in g10.items() if k.startswith("P.")})
    Pg = {k: v.cuda() for k, v in P.items()}
    if requires_grad:
        for k, v in Pg.items():
            if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
                v.requires_grad_(True)
    nerf = NeRFAll(Pg, *AABB, kernel_ptnum=5, precision="fp32", use_awp=True).train()
    assert nerf.mode == "nerf" and nerf.awpnet.input_ch == 256
    return nerf, P, Pg, g10


NERF_KW = dict(force_naive=False, return_pts0_rgb=True, retraw=True, N_samples=64, N_importance=64, perturb=0., raw_noise_std=0.,
               use_viewdirs=True, white_bkgd=False, inference=False, near=0., far=1.)
KMAT_N = torch.tensor([[400.0, 0, 200.0], [0, 400.0, 200.0], [0, 0, 1.0]])


def test_awp_on_nerf_mode_features_matches_the_reference():
    """run_nerf.py:203-212 / networks/nerf.py:140-150: AWP with input_ch = netwidth (256) on the trunk features of mode = nerf."""
    nerf, _, _, g = nerf_awp_model(False)
    with torch.no_grad():
        rgb, rgb1, _, other = nerf(400, 400, KMAT_N, chunk=32768, rays=g["rays"].cuda(), rays_info={"images_idx": g["images_idx"].cuda()}, **NERF_KW)
    assert_close(rgb, g["rgb"], "blended rgb", rtol=1e-4, atol=2e-4)
    assert_close(rgb1, g["rgb1"], "blended rgb1", rtol=1e-4, atol=2e-5)
    assert_close(other["ccw_fine"], g["ccw_fine"], "ccw_fine", rtol=1e-4, atol=2e-5)
    assert_close(other["rgb_awp"], g["rgb_awp"], "rgb_awp", rtol=1e-4, atol=2e-4)


def test_awp_on_nerf_mode_features_gradients():
    """Gradients of the AWP term: the AWP net against autograd on the oracle (its inputs taken from the CUDA forward), and the MLP
    trunk receives a gradient through depth_feature."""
    nerf, P, Pg, g = nerf_awp_model(True)
    gen = torch.Generator().manual_seed(51)
    G = torch.randn(12, 3, generator=gen)
    rgb, rgb1, _, other = nerf(400, 400, KMAT_N, chunk=32768, rays=g["rays"].cuda(), rays_info={"images_idx": g["images_idx"].cuda()}, **NERF_KW)
    (other["rgb_awp"] * G.cuda()).sum().backward()
    lr = nerf.last_render
    Po = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and k.startswith("awpnet.")) else v) for k, v in P.items()}
    rb, z = lr["ray_batch"].detach().cpu(), lr["z_vals"].detach().cpu()
    o, d, vd = rb[:, :3], rb[:, 3:6], rb[:, -3:]
    raw, feat = oc.nerf_mlpforward(P, "mlp_fine.", o[:, None] + d[:, None] * z[..., None], vd)
    rgb_s = oc.nerf_raw2outputs(raw, z, d)[0]
    emb = lr["img_embed"].detach().cpu()
    ccw = oc.awp_forward(Po, feat, z, d, emb, 5)
    ccw = ccw + ccw * 0.05
    ccw = ccw / ccw.sum(-1, keepdim=True)
    o_awp = oc.rbk_weighted_sum(rgb_s, ccw)
    assert_close(other["rgb_awp"], o_awp, "rgb_awp vs oracle", rtol=1e-4, atol=5e-5)
    names = [k for k in Po if k.startswith("awpnet.") and Po[k].is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    grads = torch.autograd.grad((o_awp * G).sum(), [Po[k] for k in names], allow_unused=True)
    checked = 0
    for k, ref in zip(names, grads):
        if ref is None or float(ref.abs().max()) == 0.0 or k.endswith("MAM.linear.bias"):
            continue
        b = ref
        scale = float(b.abs().max())
        if scale < 1e-8:          # (the softmax-over-samples logit weights here: gradients of ~1e-10, pure cancellation noise)
            continue
        assert_close(Pg[k].grad, b, "d " + k, rtol=2e-3, atol=2e-3 * scale)
        checked += 1
    assert checked >= 20, checked
    gt = Pg["mlp_fine.pts_linears.7.weight"].grad
    assert gt is not None and float(gt.abs().max()) > 0
