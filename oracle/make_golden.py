"""TEST INFRASTRUCTURE ONLY -- pins the oracle against the UNMODIFIED reference and writes tests/golden/*.npz.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
For every case the reference's own modules (imported through oracle/reference_harness.py) and the restatement in
oracle/evdeblur_oracle.py are run on the same seeded inputs; agreement is asserted here, and the reference's outputs
are stored as the golden vectors that `tests/` replay on machines without /root/reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_harness as rh  # noqa: E402
import evdeblur_oracle as oc    # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
AABB = ((-1.5, -1.5, -1.0), (1.5, 1.5, 1.0))
CFG = {"aabb_min": AABB[0], "aabb_max": AABB[1], "rmnearplane": 0}
CFG_REFNORM = dict(CFG, pdf_norm="torch_sum")   # reference's own torch.sum normaliser -> bit-level pin
H = W = 400
FOCAL = 400.0
KMAT = torch.tensor([[FOCAL, 0, 200.0], [0, FOCAL, 200.0], [0, 0, 1.0]])
COARSE_VOX, FINE_VOX = 18 * 18 * 12, 36 * 36 * 24
KW = dict(use_viewdirs=True, white_bkgd=False, inference=False, near=0., far=1.)


def close(a, b, name, rtol=1e-5, atol=1e-6):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    err = (a.double() - b.double()).abs().max().item() if a.numel() else 0.0
    ok = torch.allclose(a.double(), b.double(), rtol=rtol, atol=atol)
    print(f"  {name:28s} max|diff| = {err:.3e}  {'ok' if ok else 'MISMATCH'}")
    assert ok, name


def npy(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def ref_inds_order(weights0, z0, n_imp, u=None):
    """`inds` / `order` exactly as the reference's own expressions produce them (utils/rays.py:151-177,
    renderer.py:200-205) -- the reference discards them, so they are recomputed with its formulas."""
    w = weights0[..., 1:-1] + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        u = torch.linspace(0., 1., steps=n_imp).expand(list(cdf.shape[:-1]) + [n_imp])
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    return inds, cdf


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    E = 5
    args = rh.blurfactory_args(E=E, coarse_n_voxels=COARSE_VOX, fine_n_voxels=FINE_VOX, aabb=AABB)
    nerf, crf = rh.build_reference(args, seed=0)
    from utils.rays import sample_pdf as ref_sample_pdf
    from utils.events import egm_loss as ref_egm
    from utils import edi as ref_edi
    P = {k: v.detach().clone() for k, v in nerf.state_dict().items()}
    Pc = {k: v.detach().clone() for k, v in crf.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "params_small.npz"), **npy(P), **{"crf." + k: v for k, v in npy(Pc).items()})
    print("grid sizes", nerf.mlp_coarse.gridSize.tolist(), nerf.mlp_fine.gridSize.tolist(),
          oc.vm_grid_size(*AABB, COARSE_VOX), oc.vm_grid_size(*AABB, FINE_VOX))
    assert oc.vm_grid_size(*AABB, COARSE_VOX) == nerf.mlp_coarse.gridSize.tolist()

    # ---- case 0: BASELINE config[0]  256 rays x 1 exposure, 64 + 0 samples, forward, perturb = 0 ------------------
    print("case0: 256 rays, 64+0 samples, E=1 (render_rays)")
    rays, idx = rh.synthetic_rays(256, seed=10)
    rb = oc.build_ray_batch(H, W, FOCAL, rays)
    nerf.train()
    with torch.no_grad():
        rgb, depth, acc, ex = nerf.render(H, W, KMAT, 32768, rays, retraw=True, N_samples=64, N_importance=0,
                                          perturb=0., raw_noise_std=0., **KW)
        mine = oc.render_rays(P, CFG, rb, 64, 0)
    close(mine["rgb_map"], rgb, "rgb_map"); close(mine["depth_map"], depth, "depth_map")
    close(mine["acc_map"], acc, "acc_map"); close(mine["weights"], ex["weights"], "weights")
    np.savez_compressed(os.path.join(OUT, "case0_coarse256.npz"), rays=rays.numpy(), ray_batch=rb.numpy(),
                        rgb_map=rgb.numpy(), depth_map=depth.numpy(), acc_map=acc.numpy(),
                        weights=ex["weights"].numpy(), z_vals=ex["z_vals"].numpy())

    # ---- case 1: 48 rays x 5 exposures, 64 + 64, RBK + AWP training forward, deterministic ------------------------
    print("case1: 48 rays x 5 exposures, 64+64, RBK+AWP forward_train, perturb=0")
    N = 48
    rays, idx = rh.synthetic_rays(N, seed=11)
    with torch.no_grad():
        new_rays, weight1, _, kex = nerf.kernelsnet(H, W, KMAT, rays, {"images_idx": idx}, return_img_embed=True)
        rgb_s, depth_s, acc_s, ex = nerf.render(H, W, KMAT, 32768, new_rays.reshape(-1, 3, 2), retraw=True,
                                                N_samples=64, N_importance=64, perturb=0., raw_noise_std=0., **KW)
        rgb, rgb1, other_loss, other = nerf(H, W, KMAT, chunk=32768, rays=rays, rays_info={"images_idx": idx},
                                            force_naive=False, return_pts0_rgb=True, retraw=True, N_samples=64,
                                            N_importance=64, perturb=0., raw_noise_std=0., **KW)
        mine = oc.forward_train(P, CFG_REFNORM, H, W, FOCAL, rays, idx, E, 64, 64)
        spec = oc.forward_train(P, CFG, H, W, FOCAL, rays, idx, E, 64, 64)
    r = mine["render"]
    close(mine["new_rays"], new_rays, "rbk new_rays"); close(mine["weight1"], weight1, "rbk weight")
    for k in ("rgb_map", "depth_map", "acc_map"):
        close(r[k], {"rgb_map": rgb_s, "depth_map": depth_s, "acc_map": acc_s}[k], k)
    for k in ("rgb0", "depth0", "acc0", "weights0", "z_vals0", "z_vals", "weights", "z_std", "depth_feature"):
        close(r[k], ex[k], k, rtol=2e-5, atol=2e-6)
    close(mine["rgb"], rgb, "blended rgb"); close(mine["rgb1"], rgb1, "blended rgb1")
    close(mine["rgb_awp"], other["rgb_awp"], "rgb_awp", rtol=1e-4)
    close(mine["stage1_rgb_pts0"], other["stage1_rgb_pts0"], "stage1_rgb_pts0")
    tv = oc.tv_loss_app(P, "mlp_coarse.") + oc.tv_loss_app(P, "mlp_fine.")
    close(tv * 5, other_loss["TV"], "TV")
    # indices: reference expressions vs the oracle's specified arithmetic, on identical weights0
    inds_ref, cdf_ref = ref_inds_order(ex["weights0"], ex["z_vals0"], 64)
    z_mid = .5 * (ex["z_vals0"][..., 1:] + ex["z_vals0"][..., :-1])
    zs_ref = ref_sample_pdf(z_mid, ex["weights0"][..., 1:-1], 64, det=True)
    zs_rn, inds_rn = oc.sample_pdf(z_mid, ex["weights0"][..., 1:-1], 64, norm="torch_sum")
    assert torch.equal(inds_rn, inds_ref) and torch.equal(zs_rn, zs_ref), "oracle(torch_sum normaliser) != reference"
    print("  inds / z_samples with the reference's normaliser: bit-exact")
    zs_or, inds_or = oc.sample_pdf(z_mid, ex["weights0"][..., 1:-1], 64)
    mism = (inds_ref != inds_or)
    print(f"  specified (fp64) normaliser vs reference: inds mismatches {int(mism.sum())} / {mism.numel()} "
          f"(last column u=1: {int(mism[:, -1].sum())}), max|dz| = {(zs_or - zs_ref).abs().max().item():.3e}")
    for k in ("rgb_map", "depth_map", "acc_map", "z_vals", "weights"):
        close(spec["render"][k], ex[k if k in ex else k] if k in ex else {"rgb_map": rgb_s, "depth_map": depth_s, "acc_map": acc_s}[k],
              "spec-norm " + k, rtol=1e-4, atol=2e-4)
    z_all_ref, order_ref = torch.sort(torch.cat([ex["z_vals0"], zs_ref], -1), -1)
    z_all_or, order_or = oc.merge_samples(ex["z_vals0"], zs_ref)
    assert torch.equal(z_all_ref, z_all_or)
    print(f"  order mismatches (ties only): {int((order_ref != order_or).sum())}")
    np.savez_compressed(
        os.path.join(OUT, "case1_train48x5.npz"), rays=rays.numpy(), images_idx=idx.numpy(),
        new_rays=new_rays.numpy(), weight1=weight1.numpy(), img_embed=kex["img_embed"].numpy(),
        rgb_map=rgb_s.numpy(), depth_map=depth_s.numpy(), acc_map=acc_s.numpy(),
        **{k: ex[k].numpy() for k in ("rgb0", "depth0", "acc0", "weights0", "z_vals0", "z_vals", "weights", "z_std")},
        depth_feature_sum=ex["depth_feature"].double().sum(-1).float().numpy(),
        depth_feature_head=ex["depth_feature"][:8].numpy(),
        inds=inds_ref.numpy(), order=order_ref.numpy(), z_samples=zs_ref.numpy(), cdf=cdf_ref.numpy(),
        rgb=rgb.numpy(), rgb1=rgb1.numpy(), rgb_awp=other["rgb_awp"].numpy(), ccw_fine=mine["ccw_fine"].numpy(),
        stage1_rgb_pts0=other["stage1_rgb_pts0"].numpy(), stage1_rgb1_pts0=other["stage1_rgb1_pts0"].numpy(),
        TV=other_loss["TV"].numpy())

    # ---- case 2: injected randomness (perturb = 1, raw_noise_std = 1), 32 rays x 1 exposure ------------------------
    print("case2: 32 rays, 64+64, perturb=1 raw_noise_std=1 with replayed RNG draws")
    rays, idx = rh.synthetic_rays(32, seed=12)
    rb = oc.build_ray_batch(H, W, FOCAL, rays)
    R = rb.shape[0]
    torch.manual_seed(1234)
    with torch.no_grad():
        rgb, depth, acc, ex = nerf.render(H, W, KMAT, 32768, rays, retraw=True, N_samples=64, N_importance=64,
                                          perturb=1., raw_noise_std=1., **KW)
    torch.manual_seed(1234)   # replay the reference's draw order: renderer.py:176, voxnerf.py:175, rays.py:162, voxnerf.py:175
    rand = {"t_rand": torch.rand(R, 64), "noise0": torch.randn(R, 63), "u": torch.rand(R, 64),
            "noise1": torch.randn(R, 127)}
    with torch.no_grad():
        mine = oc.render_rays(P, CFG_REFNORM, rb, 64, 64, perturb=1., rand=rand)
    for k, v in (("rgb_map", rgb), ("depth_map", depth), ("acc_map", acc), ("rgb0", ex["rgb0"]),
                 ("weights0", ex["weights0"]), ("z_vals", ex["z_vals"]), ("weights", ex["weights"])):
        close(mine[k], v, k, rtol=2e-5, atol=2e-6)
    inds_ref, _ = ref_inds_order(ex["weights0"], ex["z_vals0"], 64, rand["u"])
    print(f"  inds mismatches: {int((inds_ref != mine['inds']).sum())} / {inds_ref.numel()}")
    np.savez_compressed(os.path.join(OUT, "case2_perturb32.npz"), ray_batch=rb.numpy(), **npy(rand),
                        rgb_map=rgb.numpy(), depth_map=depth.numpy(), acc_map=acc.numpy(), rgb0=ex["rgb0"].numpy(),
                        weights0=ex["weights0"].numpy(), z_vals0=ex["z_vals0"].numpy(), z_vals=ex["z_vals"].numpy(),
                        weights=ex["weights"].numpy(), inds=inds_ref.numpy())

    # ---- case 3: loss heads: CRF, luma, EGM loss, MSE ---------------------------------------------------------------
    print("case3: CRF / EGM / MSE")
    g = torch.Generator().manual_seed(13)
    M = 96
    x0, x1 = torch.rand(M, 3, generator=g) * 0.9 + 0.05, torch.rand(M, 3, generator=g) * 0.9 + 0.05
    pol = torch.stack([-(torch.rand(M, generator=g) < 0.5).float() * 2, (torch.rand(M, generator=g) < 0.5).float()], -1)
    target = torch.rand(M, 3, generator=g)
    cmask = torch.nn.functional.one_hot(torch.randint(0, 3, (M,), generator=g), 3).bool()
    bii = (torch.tensor([0.2, 0.2]) * pol).sum(-1)
    with torch.no_grad():
        enc = crf(x0, mode="encode_rgb", skip_learn_crf=False)
        l0 = crf(x0, mode="encode_luma", skip_learn_crf=False, ev_extra_feat=pol)
        l1 = crf(x1, mode="encode_luma", skip_learn_crf=False, ev_extra_feat=pol)
        l0s = crf(x0, mode="encode_luma", skip_learn_crf=True, ev_extra_feat=pol)
        cpol = torch.zeros(M, 3, 2); cpol[cmask] = pol
        c0 = crf(x0, mode="encode_luma", skip_learn_crf=False, ev_extra_feat=cpol, tonemap_only=True)
        c1 = crf(x1, mode="encode_luma", skip_learn_crf=False, ev_extra_feat=cpol, tonemap_only=True)
        l0n = crf(x0, mode="encode_luma", skip_learn_crf=False)   # zero-padded extra feats, tonemapping.py:83-86
        e_gray = ref_egm(l0, l1, bii)
        e_col = ref_egm(c0, c1, bii, color_mask=cmask, color_weight=[0.4, 0.2, 0.4])
        mse = torch.mean((enc - target) ** 2)
    close(oc.encode_rgb(Pc, x0), enc, "encode_rgb")
    close(oc.encode_luma(Pc, x0, ev_extra_feat=pol), l0, "encode_luma")
    close(oc.encode_luma(Pc, x0, ev_extra_feat=pol, skip_learn=True), l0s, "encode_luma skip")
    close(oc.encode_luma(Pc, x0, ev_extra_feat=cpol, tonemap_only=True), c0, "encode_luma color")
    close(oc.encode_luma(Pc, x0), l0n, "encode_luma nofeat")
    close(oc.egm_loss(oc.encode_luma(Pc, x0, ev_extra_feat=pol), oc.encode_luma(Pc, x1, ev_extra_feat=pol), bii), e_gray, "egm gray")
    close(oc.egm_loss(oc.encode_luma(Pc, x0, ev_extra_feat=cpol, tonemap_only=True),
                      oc.encode_luma(Pc, x1, ev_extra_feat=cpol, tonemap_only=True), bii, cmask, [0.4, 0.2, 0.4]),
          e_col, "egm color")
    close(oc.img2mse(oc.encode_rgb(Pc, x0), target), mse, "img2mse")
    np.savez_compressed(os.path.join(OUT, "case3_loss.npz"), x0=x0.numpy(), x1=x1.numpy(), pol=pol.numpy(),
                        cpol=cpol.numpy(), cmask=cmask.numpy(), bii=bii.numpy(), target=target.numpy(),
                        enc=enc.numpy(), l0=l0.numpy(), l1=l1.numpy(), l0s=l0s.numpy(), c0=c0.numpy(), c1=c1.numpy(),
                        l0n=l0n.numpy(), e_gray=e_gray.numpy(), e_col=e_col.numpy(), mse=mse.numpy())

    # ---- case 4: EDI prior ------------------------------------------------------------------------------------------
    print("case4: EDI")
    rng = np.random.default_rng(14)
    w, h, n_ev = 40, 32, 6000
    ev_x = rng.uniform(-0.4, w - 0.2, n_ev).astype(np.float32)
    ev_y = rng.uniform(0, h - 0.2, n_ev).astype(np.float32)
    ev_x[::7] = np.round(ev_x[::7]); ev_y[::5] = np.round(ev_y[::5])     # integer coordinates: duplicate-tap rule
    ev_x[ev_x < -0.3] = np.clip(ev_x[ev_x < -0.3], -0.9, None)           # a few negative x: np.add.at wrap quirk
    ev_p = rng.integers(0, 2, n_ev).astype(np.float32)
    ev_t = np.sort(rng.uniform(1000., 2000., n_ev))
    blurry = rng.uniform(0.05, 1.0, (h, w, 3)).astype(np.float32)
    t0, t1, steps, cpos, cneg = 1100.0, 1900.0, 9, 0.2, 0.25
    ts = np.linspace(t0, t1, steps)
    il, ir = np.searchsorted(ev_t, ts), np.searchsorted(ev_t, ts, side="right")
    biis = []
    for j in range(steps - 1):
        a, b = il[j], ir[j + 1]
        bi = ref_edi.brightness_increment_image(ev_x[a:b], ev_y[a:b], ev_p[a:b], w, h, cpos, cneg, interpolate=True)
        biis.append(bi[..., None].repeat(3, axis=-1))
    biis = np.stack(biis, 0)
    sharp = ref_edi.deblur_double_integral(blurry, biis)
    mine_sharp = oc.edi_prior_image(ev_x, ev_y, ev_t, ev_p, blurry, t0, t1, w, h, cpos, cneg, steps)
    close(mine_sharp, sharp, "edi sharp"); close(oc.edi_inner(biis), ref_edi.inner_double_integral(biis), "edi inner")
    np.savez_compressed(os.path.join(OUT, "case4_edi.npz"), ev_x=ev_x, ev_y=ev_y, ev_p=ev_p, ev_t=ev_t, blurry=blurry,
                        t0=t0, t1=t1, steps=steps, cpos=cpos, cneg=cneg, bii=biis[..., 0], sharp=sharp)

    # ---- case 5: mode = nerf (run_network), separate small parameter set --------------------------------------------
    print("case5: mode=nerf 8x256, 24 rays, 64+64")
    args_n = rh.blurfactory_args(E=E, mode="nerf", use_awp=True, rgb_add_bias=True)
    nerf_n, _ = rh.build_reference(args_n, seed=5)
    Pn = {k: v.detach().clone() for k, v in nerf_n.state_dict().items() if k.startswith("mlp_")}
    rays, idx = rh.synthetic_rays(24, seed=15)
    rb = oc.build_ray_batch(H, W, FOCAL, rays)
    nerf_n.train()
    with torch.no_grad():
        rgb, depth, acc, ex = nerf_n.render(H, W, KMAT, 32768, rays, retraw=True, N_samples=64, N_importance=64,
                                            perturb=0., raw_noise_std=0., **KW)
        o, d, vd = rb[:, :3], rb[:, 3:6], rb[:, -3:]
        pts = o[:, None] + d[:, None] * ex["z_vals"][..., None]
        raw_ref, feat_ref = nerf_n.mlp_fine.mlpforward(pts, vd, nerf_n.embed_fn, nerf_n.embeddirs_fn)
        raw, feat = oc.nerf_mlpforward(Pn, "mlp_fine.", pts, vd)
        out = oc.nerf_raw2outputs(raw, ex["z_vals"], d)
    close(raw, raw_ref, "nerf raw"); close(feat, feat_ref, "nerf feature")
    close(out[0], rgb, "nerf rgb_map"); close(out[3], ex["weights"], "nerf weights"); close(out[4], depth, "nerf depth")
    np.savez_compressed(os.path.join(OUT, "case5_nerf24.npz"), ray_batch=rb.numpy(), z_vals=ex["z_vals"].numpy(),
                        raw=raw_ref.numpy(), feature_sum=feat_ref.double().sum(-1).float().numpy(), rgb_map=rgb.numpy(),
                        depth_map=depth.numpy(), acc_map=acc.numpy(), weights=ex["weights"].numpy(),
                        **{"P." + k: v.numpy() for k, v in Pn.items() if k.startswith("mlp_fine.")})

    # ---- case 6: grid_sample known answers vs explicit taps (incl. out-of-range points) -----------------------------
    print("case6: VM sample known answers")
    g = torch.Generator().manual_seed(16)
    pts = (torch.rand(40, 8, 3, generator=g) * 2 - 1) * torch.tensor([1.7, 1.7, 1.2])    # ~12% outside the aabb
    pts[0, 0] = torch.tensor(AABB[0]); pts[0, 1] = torch.tensor(AABB[1]); pts[0, 2] = torch.tensor([0., 0., 0.])
    with torch.no_grad():
        fc_ref, ff_ref = nerf.mlp_coarse.sample(pts), nerf.mlp_fine.sample(pts)
    close(oc.vm_sample(P, "mlp_coarse.", pts, *AABB), fc_ref, "vm coarse (grid_sample)")
    close(oc.vm_sample_taps(P, "mlp_coarse.", pts, *AABB), fc_ref, "vm coarse (taps)", rtol=1e-4, atol=1e-6)
    close(oc.vm_sample_taps(P, "mlp_fine.", pts, *AABB), ff_ref, "vm fine (taps)", rtol=1e-4, atol=1e-6)
    np.savez_compressed(os.path.join(OUT, "case6_vm.npz"), pts=pts.numpy(), ft_coarse=fc_ref.numpy(), ft_fine=ff_ref.numpy())
    sizes = {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))}
    print("fixtures:", sizes, "total", sum(sizes.values()))


if __name__ == "__main__":
    main()
