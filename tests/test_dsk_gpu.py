"""GPU parity of the deformable sparse kernel (DSK) path through the C ABI: edn_dsk_rays_fwd / edn_dsk_rays_bwd against vectors
written by the UNMODIFIED reference BlurModel (forward outputs and autograd parameter gradients, tests/golden/case9_dsk.npz),
edn_build_ray_batch_bwd against autograd on the oracle, and the NeRFAll facade with kernel_type = DSK end to end (forward against
the reference NeRFAll's own outputs, gradients against autograd on the oracle).

Tolerances: forward 1e-5 relative (+1e-6 abs; new_rays of order 1); gradients within 1e-4 of each tensor's max magnitude (the
reductions over rays are cuBLAS GEMMs and atomics in a different order than autograd's); end to end the usual 1e-4 / 2e-4 of the
render path and 5e-3 of max for the kernel-net tensors (per-ray terms cancel, as for RBK)."""
import pytest
import torch

import evdeblur_oracle as oc
from test_dsk_cpu import DSK_CFG, KMAT, H, W, case
from util import AABB, CFG, assert_close, golden, small_params

pytestmark = pytest.mark.gpu


def grad_close(a, b, name, tol=1e-4):
    b = torch.as_tensor(b)
    scale = float(b.abs().max())
    assert scale > 0, f"{name}: reference gradient is identically zero (test is vacuous)"
    assert_close(a, b, name, rtol=tol, atol=tol * scale)


def model_for(name, P, requires_grad=False):
    from evdeblurnerf_b200 import BlurModel
    c = DSK_CFG[name]
    Pg = {k: v.cuda() for k, v in P.items()}
    if requires_grad:
        for v in Pg.values():
            if v.is_floating_point():
                v.requires_grad_(True)
    return BlurModel(Pg, c["num_pt"], kernel_hwindow=c["kernel_hwindow"], in_embed=c["in_embed"], spatial_embed=c["spatial_embed"],
                     random_hwindow=0.0), Pg


def info_of(io):
    return {"rays_x": io["rays_x"].cuda(), "rays_y": io["rays_y"].cuda(), "images_idx": io["images_idx"].cuda(), "poses": io["poses"].cuda()}


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_dsk_forward_matches_the_reference(name):
    P, _, io = case(golden("case9_dsk"), name)
    m, _ = model_for(name, P)
    noise = io["noise"].cuda() if "noise" in io else None
    new_rays, weight, align, extras = m(H, W, KMAT, None, info_of(io), return_img_embed=True, noise=noise)
    assert_close(new_rays, io["new_rays"], "new_rays", rtol=1e-5, atol=1e-6)
    assert_close(weight, io["weight"], "weight", rtol=1e-5, atol=1e-7)
    assert_close(align.reshape(1), io["align"], "align", rtol=1e-5, atol=1e-7)
    assert_close(extras["img_embed"], P["kernelsnet.img_embed.img_embed"][io["images_idx"].reshape(-1)], "img_embed", rtol=0, atol=0)


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_dsk_backward_matches_reference_autograd(name):
    P, G, io = case(golden("case9_dsk"), name)
    m, Pg = model_for(name, P, requires_grad=True)
    noise = io["noise"].cuda() if "noise" in io else None
    new_rays, weight, align, _ = m(H, W, KMAT, None, info_of(io), noise=noise)
    loss = (new_rays * io["G_rays"].cuda()).sum() + (weight * io["G_w"].cuda()).sum() + align * float(io["g_align"][0])
    loss.backward()
    checked = 0
    for k, ref in G.items():
        got = Pg["kernelsnet." + k].grad
        if float(ref.abs().max()) == 0.0:      # a parameter the option set does not reach (none today)
            assert got is None or float(got.abs().max()) == 0.0, k
            continue
        assert got is not None, k
        grad_close(got, ref, "d " + k)
        checked += 1
    assert checked >= 8


def test_dsk_noise_is_drawn_when_random_hwindow_is_set():
    from evdeblurnerf_b200 import BlurModel
    P, _, io = case(golden("case9_dsk"), "a")
    Pg = {k: v.cuda() for k, v in P.items()}
    m = BlurModel(Pg, 5, in_embed=3, random_hwindow=0.25)
    a = m(H, W, KMAT, None, info_of(io))[0]
    b = m(H, W, KMAT, None, info_of(io))[0]
    assert float((a - b).abs().max()) > 0          # blurmodel.py:125-127: a fresh randn per call
    with pytest.raises(RuntimeError):
        BlurModel(Pg, 5, in_embed=2)                # linears.0's input width does not match
    with pytest.raises(NotImplementedError):
        m(H, W, KMAT, None, info_of(io), feats=torch.zeros(1))      # PBE features


@pytest.mark.parametrize("ndc", [True, False])
def test_build_ray_batch_backward_matches_autograd(ndc):
    from evdeblurnerf_b200 import _lib
    from evdeblurnerf_b200._lib import check, ptr, stream_ptr
    from util import synthetic_rays
    rays, _ = synthetic_rays(257, seed=5)
    g = torch.Generator().manual_seed(6)
    d_rb = torch.randn(257, 11, generator=g)
    rg = rays.clone().requires_grad_(True)
    rb = oc.build_ray_batch(H, W, 400.0, rg, ndc=ndc)
    (rb * d_rb).sum().backward()
    out = torch.empty(257, 3, 2, device="cuda")
    rays_d, d_rb_d = rays.cuda().contiguous(), d_rb.cuda().contiguous()       # keep the device copies alive across the launch
    check(_lib.load().edn_build_ray_batch_bwd(ptr(rays_d), 257, H, W, 400.0, 1 if ndc else 0, ptr(d_rb_d), ptr(out), stream_ptr()),
          "edn_build_ray_batch_bwd")
    grad_close(out, rg.grad, "d rays")


def e2e_model(requires_grad):
    from evdeblurnerf_b200 import NeRFAll
    g = golden("case9_dsk")
    P, _ = small_params()
    P = {k: v for k, v in P.items() if not k.startswith(("kernelsnet.", "awpnet."))}
    P.update({"kernelsnet." + k[len("e2e.param."):]: v for k, v in g.items() if k.startswith("e2e.param.")})
    Pg = {k: v.cuda() for k, v in P.items()}
    if requires_grad:
        for v in Pg.values():
            if v.is_floating_point():
                v.requires_grad_(True)
    nerf = NeRFAll(Pg, *AABB, kernel_ptnum=5, precision="fp32", kernel_cfg=dict(in_embed=3, spatial_embed=0, kernel_hwindow=10, random_hwindow=0.0))
    assert nerf.kernel_type == "DSK"
    io = {k[4:]: v for k, v in g.items() if k.startswith("e2e.") and not k.startswith("e2e.param.")}
    return nerf.train(), P, Pg, io


KW = dict(force_naive=False, return_pts0_rgb=True, retraw=True, N_samples=64, N_importance=64, perturb=0., raw_noise_std=0.,
          use_viewdirs=True, white_bkgd=False, inference=False, near=0., far=1.)


def test_nerfall_dsk_forward_matches_the_reference():
    nerf, _, _, io = e2e_model(False)
    with torch.no_grad():
        rgb, rgb1, other_loss, other = nerf(H, W, KMAT, chunk=32768, rays=torch.zeros(24, 3, 2).cuda(), rays_info=info_of(io), **KW)
    assert_close(rgb, io["rgb"], "blended rgb", rtol=1e-4, atol=2e-4)
    assert_close(rgb1, io["rgb1"], "blended rgb1", rtol=1e-4, atol=2e-5)
    assert_close(other["stage1_rgb_pts0"], io["stage1_rgb_pts0"], "stage1_rgb_pts0", rtol=1e-4, atol=2e-4)
    assert_close(other_loss["align"].reshape(1), io["align"], "align", rtol=1e-5, atol=1e-7)
    assert_close(other_loss["TV"].reshape(1), io["TV"], "TV", rtol=1e-5)


def test_nerfall_dsk_gradients_match_autograd_on_the_oracle():
    nerf, P, Pg, io = e2e_model(True)
    gen = torch.Generator().manual_seed(31)
    G, G1 = torch.randn(24, 3, generator=gen), torch.randn(24, 3, generator=gen)
    rgb, rgb1, other_loss, _ = nerf(H, W, KMAT, chunk=32768, rays=torch.zeros(24, 3, 2).cuda(), rays_info=info_of(io), **KW)
    loss = (rgb * G.cuda()).sum() + (rgb1 * G1.cuda()).sum() + 0.3 * other_loss["align"].sum()
    loss.backward()
    # oracle: same functional, fine pass evaluated at the depths the CUDA sampler produced (the sampler is ill-conditioned in the
    # last bits of weights0; z_samples carry no gradient in the reference either, renderer.py:203)
    from util import oracle_fine_at
    Po = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
    cfg = dict(num_pt=5, kernel_hwindow=10, in_embed=3, spatial_embed=0, num_hidden=3, short_cut=False, isglobal=False, optim_trans=False,
               optim_sv_trans=False)
    new_rays, weight, align = oc.dsk_forward(Po, cfg, H, W, KMAT, io["rays_x"], io["rays_y"], io["images_idx"], io["poses"])
    rb = oc.build_ray_batch(H, W, 400.0, new_rays.reshape(-1, 3, 2))
    c = oc.render_rays(Po, CFG, rb, 64, 0)
    f = oracle_fine_at(Po, rb, nerf.last_render["z_vals"].cpu(), None)
    o_rgb = torch.sum(f["rgb_map"].reshape(24, 5, 3) * weight[..., None], 1)
    o_rgb1 = torch.sum(c["rgb_map"].reshape(24, 5, 3) * weight[..., None], 1)
    assert_close(rgb, o_rgb, "rgb vs oracle at the same depths", rtol=1e-4, atol=2e-5)
    oloss = (o_rgb * G).sum() + (o_rgb1 * G1).sum() + 0.3 * align
    names = [k for k in Po if k.startswith("kernelsnet.") and Po[k].is_floating_point()]
    grads = torch.autograd.grad(oloss, [Po[k] for k in names], allow_unused=True)
    checked = 0
    for k, ref in zip(names, grads):
        if ref is None or float(ref.abs().max()) == 0.0:
            continue
        grad_close(Pg[k].grad, ref, "d " + k, tol=5e-3)
        checked += 1
    assert checked >= 8
    assert Pg["mlp_fine.app_plane.0"].grad is not None and float(Pg["mlp_fine.app_plane.0"].grad.abs().max()) > 0


def test_trainer_steps_with_a_dsk_kernel_and_the_align_term():
    """Trainer.step with kernel_type = DSK: the loss includes kernel_align_weight * align (run_nerf.py:502-504), every DSK tensor
    moves, the loss of a fixed batch falls."""
    from evdeblurnerf_b200.trainer import Trainer
    g = golden("case9_dsk")
    P, _ = small_params()
    P = {k: v for k, v in P.items() if not k.startswith(("kernelsnet.", "awpnet."))}
    P.update({"kernelsnet." + k[len("e2e.param."):]: v for k, v in g.items() if k.startswith("e2e.param.")})
    io = {k[4:]: v for k, v in g.items() if k.startswith("e2e.") and not k.startswith("e2e.param.")}
    tr = Trainer({k: v.cuda() for k, v in P.items()}, None, *AABB, kernel_ptnum=5, precision="fp32", tv_loss_weight=0.0, lrate=2e-3,
                 render_kwargs=dict(N_samples=32, N_importance=32, perturb=0., raw_noise_std=0.),
                 kernel_cfg=dict(in_embed=3, random_hwindow=0.0), schedule=dict(kernel_align_weight=0.1), check_numerics_every=0)
    assert tr.nerf.kernel_type == "DSK" and not tr.fuse_event_renders
    before = {k: v.detach().clone() for k, v in tr.nerf.params.items() if k.startswith("kernelsnet.")}
    batch = dict(info_of(io), rays=torch.zeros(24, 3, 2).cuda(), rgbsf=torch.full((24, 3), 0.25).cuda())
    losses = [float(tr.step(batch, H, W, KMAT)["loss"]) for _ in range(12)]
    assert losses[-1] < losses[0], losses
    moved = [k for k, v in before.items() if float((tr.nerf.params[k].detach() - v).abs().max()) > 0]
    assert len(moved) == len(before), sorted(set(before) - set(moved))


def awp_model(requires_grad):
    from evdeblurnerf_b200 import NeRFAll
    g = golden("case9_dsk")
    P, _ = small_params()
    P = {k: v for k, v in P.items() if not k.startswith("kernelsnet.")}
    P.update({"kernelsnet." + k[len("e2e.param."):]: v for k, v in g.items() if k.startswith("e2e.param.")})
    Pg = {k: v.cuda() for k, v in P.items()}
    if requires_grad:
        for k, v in Pg.items():
            if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
                v.requires_grad_(True)
    nerf = NeRFAll(Pg, *AABB, kernel_ptnum=5, precision="fp32", use_awp=True,
                   kernel_cfg=dict(in_embed=3, spatial_embed=0, kernel_hwindow=10, random_hwindow=0.0)).train()
    io = {k[4:]: v for k, v in g.items() if k.startswith("e2e.") and not k.startswith("e2e.param.")}
    return nerf, P, Pg, io


def test_nerfall_dsk_with_awp_matches_the_reference():
    """renderer.py:310-343 with a non-RBK kernel: AWP's weights over the DSK points -> rgb_awp (reference NeRFAll output)."""
    nerf, _, _, io = awp_model(False)
    with torch.no_grad():
        rgb, rgb1, _, other = nerf(H, W, KMAT, chunk=32768, rays=torch.zeros(24, 3, 2).cuda(), rays_info=info_of(io), **KW)
    assert_close(rgb, io["awp.rgb"], "blended rgb", rtol=1e-4, atol=2e-4)
    assert_close(rgb1, io["awp.rgb1"], "blended rgb1", rtol=1e-4, atol=2e-5)
    assert_close(other["ccw_fine"], io["awp.ccw_fine"], "ccw_fine", rtol=1e-4, atol=1e-5)
    assert_close(other["rgb_awp"], io["awp.rgb_awp"], "rgb_awp", rtol=1e-4, atol=2e-4)


def test_nerfall_dsk_with_awp_gradients_reach_every_tensor():
    """The AWP term's gradient reaches the AWP net, the DSK view latents (through the row gather) and the fields; compared against
    autograd on the oracle for the kernel-net and AWP tensors."""
    nerf, P, Pg, io = awp_model(True)
    gen = torch.Generator().manual_seed(41)
    G = torch.randn(24, 3, generator=gen)
    rgb, rgb1, _, other = nerf(H, W, KMAT, chunk=32768, rays=torch.zeros(24, 3, 2).cuda(), rays_info=info_of(io), **KW)
    (other["rgb_awp"] * G.cuda()).sum().backward()
    from util import oracle_fine_at
    Po = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
    cfg = dict(num_pt=5, kernel_hwindow=10, in_embed=3, spatial_embed=0, num_hidden=3, short_cut=False, isglobal=False, optim_trans=False,
               optim_sv_trans=False)
    new_rays, weight, _ = oc.dsk_forward(Po, cfg, H, W, KMAT, io["rays_x"], io["rays_y"], io["images_idx"], io["poses"])
    rb = oc.build_ray_batch(H, W, 400.0, new_rays.reshape(-1, 3, 2))
    z = nerf.last_render["z_vals"].cpu()
    f = oracle_fine_at(Po, rb, z, None)
    emb = Po["kernelsnet.img_embed.img_embed"][io["images_idx"].reshape(-1).long()]
    ccw = oc.awp_forward(Po, f["depth_feature"], z, rb[:, 3:6], emb, 5)
    ccw = ccw + ccw * 0.05
    ccw = ccw / ccw.sum(-1, keepdim=True)
    o_awp = torch.sum(f["rgb_map"].reshape(24, 5, 3) * ccw[..., None], 1)
    assert_close(other["rgb_awp"], o_awp, "rgb_awp vs oracle at the same depths", rtol=1e-4, atol=5e-5)
    names = [k for k in Po if k.startswith(("kernelsnet.", "awpnet.")) and Po[k].is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    grads = torch.autograd.grad((o_awp * G).sum(), [Po[k] for k in names], allow_unused=True)
    checked = 0
    for k, ref in zip(names, grads):
        if ref is None or float(ref.abs().max()) == 0.0 or k.endswith("MAM.linear.bias"):     # (BatchNorm cancels that bias: rounding noise)
            continue
        grad_close(Pg[k].grad, ref, "d " + k, tol=5e-3)
        checked += 1
    assert checked >= 25, checked
