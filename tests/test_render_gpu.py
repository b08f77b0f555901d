"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the committed golden vectors.
Tolerance: 1e-4 relative (north_star, fp32 parity mode); sample indices bit-exact."""
import pytest
import torch

import evdeblur_oracle as oc
from util import (AABB, CFG, FOCAL, H, W, assert_close, golden, oracle_fine_at, random_params, small_params,
                  synthetic_rays)

# end-to-end tolerance on merged depths / per-sample weights: the inverse-CDF step amplifies last-bit differences of
# weights0 by up to 1/denom (denom >= 1e-5) x bin width (1/64), see util.oracle_fine_at
Z_ATOL = 5e-4

pytestmark = pytest.mark.gpu


# every test of this file runs in BOTH parity-grade precisions: "fp32" (SIMT kernels) and "tc32" (the fine pass's GEMMs on tcgen05
# with bf16 x 3 split operands, fp32 TMEM accumulation) -- same tolerances
@pytest.fixture(scope="module", params=["fp32", "tc32"])
def engine(request):
    from evdeblurnerf_b200 import RenderEngine
    P, _ = small_params()
    return RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision=request.param)


def test_vm_sample_known_answers(engine):
    g = golden("case6_vm")
    pts = g["pts"].cuda()
    assert_close(engine.vm_sample(pts, "coarse"), g["ft_coarse"], "ft_coarse", rtol=1e-4, atol=1e-6)
    assert_close(engine.vm_sample(pts, "fine"), g["ft_fine"], "ft_fine", rtol=1e-4, atol=1e-6)


def test_vm_sample_out_of_range_is_zero_padded(engine):
    P, _ = small_params()
    pts = torch.tensor([[5.0, 5.0, 5.0], [-1.5, -1.5, -1.0], [1.5, 1.5, 1.0], [1.5001, 0.0, 0.0], [0.0, -1.6, 0.99]])[None]
    ref = oc.vm_sample(P, "mlp_fine.", pts, *AABB)
    assert_close(engine.vm_sample(pts.cuda(), "fine"), ref, "oob", rtol=1e-4, atol=1e-6)
    assert float(engine.vm_sample(pts.cuda(), "fine")[0, 0].abs().max()) == 0.0


def test_case0_coarse_only(engine):
    g = golden("case0_coarse256")
    out = engine.render_rays(g["ray_batch"].cuda(), 64, retraw=True)
    assert torch.equal(out["z_vals"].cpu(), g["z_vals"])          # sample placement is bit-exact
    for k in ("rgb_map", "depth_map", "acc_map", "weights"):
        assert_close(out[k], g[k], k)


def test_sample_pdf_indices_bit_exact(engine):
    g = golden("case1_train48x5")
    z0, w0 = g["z_vals0"], g["weights0"]
    m = engine.sample_pdf_merge(z0.cuda(), w0.cuda(), 64)
    z_mid = .5 * (z0[..., 1:] + z0[..., :-1])
    zs, inds = oc.sample_pdf(z_mid, w0[..., 1:-1], 64)
    assert torch.equal(m["inds"].cpu(), inds)
    assert torch.equal(m["z_samples"].cpu(), zs)
    zv, order = oc.merge_samples(z0, zs)
    assert torch.equal(m["z_vals"].cpu(), zv)
    assert torch.equal(m["order"].cpu(), order)
    assert_close(m["z_std"], torch.std(zs, dim=-1, unbiased=False), "z_std", rtol=1e-5)
    # against the reference's own indices: only the u = 1 column may differ
    mism = m["inds"].cpu() != g["inds"]
    assert int(mism[:, :-1].sum()) == 0


def test_sample_pdf_random_u_and_ties(engine):
    g = torch.Generator().manual_seed(3)
    R = 300
    z0 = torch.sort(torch.rand(R, 64, generator=g), -1)[0]
    w0 = torch.rand(R, 64, generator=g) ** 8
    w0[:7] = 0.0                              # all-zero weights -> uniform pdf, denom < 1e-5 branch
    u = torch.rand(R, 64, generator=g)
    u[:, 0] = 0.0
    z0[5, 10:14] = z0[5, 10]                  # duplicate coarse depths -> ties
    m = engine.sample_pdf_merge(z0.cuda(), w0.cuda(), 64, u=u.cuda())
    z_mid = .5 * (z0[..., 1:] + z0[..., :-1])
    zs, inds = oc.sample_pdf(z_mid, w0[..., 1:-1], 64, u=u)
    assert torch.equal(m["inds"].cpu(), inds)
    assert torch.equal(m["z_samples"].cpu(), zs)
    zv, order = oc.merge_samples(z0, zs)
    assert torch.equal(m["z_vals"].cpu(), zv) and torch.equal(m["order"].cpu(), order)


@pytest.mark.parametrize("nc,ni", [(32, 32), (64, 64), (96, 96), (64, 17)])
def test_sample_pdf_ragged_sizes(engine, nc, ni):
    g = torch.Generator().manual_seed(nc * 100 + ni)
    R = 33
    z0 = torch.sort(torch.rand(R, nc, generator=g), -1)[0]
    w0 = torch.rand(R, nc, generator=g)
    m = engine.sample_pdf_merge(z0.cuda(), w0.cuda(), ni)
    zs, inds = oc.sample_pdf(.5 * (z0[..., 1:] + z0[..., :-1]), w0[..., 1:-1], ni)
    assert torch.equal(m["inds"].cpu(), inds) and torch.equal(m["z_samples"].cpu(), zs)
    zv, order = oc.merge_samples(z0, zs)
    assert torch.equal(m["z_vals"].cpu(), zv) and torch.equal(m["order"].cpu(), order)


@pytest.mark.parametrize("nc,ni", [(64, 64), (517, 200), (1024, 1024)])
def test_sample_pdf_wide_dynamic_range_stays_bit_exact(engine, nc, ni):
    """The sampler's warp-parallel fp64 scan must return the bits of the sequential fp64 cumsum the oracle (torch CPU) runs: weights
    spanning 0 .. 1 with denormal-small, 1e-30, 1e-7-sized and O(1) entries, long rows, ragged tails (nc - 2 not a multiple of 32)."""
    g = torch.Generator().manual_seed(nc + ni)
    R = 64
    z0 = torch.sort(torch.rand(R, nc, generator=g), -1)[0]
    w0 = torch.rand(R, nc, generator=g)
    scale = torch.tensor([0.0, 1e-38, 1e-30, 1e-7, 1e-3, 1.0])[torch.randint(0, 6, (R, nc), generator=g)]
    w0 = w0 * scale
    w0[0] = 1.0                                # all-equal row
    w0[1, ::2] = 0.0
    u = torch.rand(R, ni, generator=g)
    m = engine.sample_pdf_merge(z0.cuda(), w0.cuda(), ni, u=u.cuda())
    zs, inds = oc.sample_pdf(.5 * (z0[..., 1:] + z0[..., :-1]), w0[..., 1:-1], ni, u=u)
    assert torch.equal(m["inds"].cpu(), inds)
    assert torch.equal(m["z_samples"].cpu(), zs)


def test_case1_c2f_render(engine):
    P, _ = small_params()
    g = golden("case1_train48x5")
    rb = oc.build_ray_batch(H, W, FOCAL, g["new_rays"].reshape(-1, 3, 2))
    out = engine.render_rays(rb.cuda(), 64, retraw=True, N_importance=64, use_awp=True, want_indices=True)
    ref = oc.render_rays(P, CFG, rb, 64, 64, want_feature=True)
    assert torch.equal(out["z_vals0"].cpu(), ref["z_vals0"])
    for k in ("rgb0", "depth0", "acc0", "weights0"):
        assert_close(out[k], ref[k], k)
        assert_close(out[k], g[k], "golden " + k)
    # indices: the CUDA weights0 differ from the oracle's in the last bits, so compare on the CUDA weights themselves
    z0, w0 = out["z_vals0"].cpu(), out["weights0"].cpu()
    zs, inds = oc.sample_pdf(.5 * (z0[..., 1:] + z0[..., :-1]), w0[..., 1:-1], 64)
    assert torch.equal(out["inds"].cpu(), inds)
    # fine pass, tight, at the depths the CUDA sampler produced
    fin = oracle_fine_at(P, rb, out["z_vals"].cpu())
    for k in ("weights", "rgb_map", "depth_map", "acc_map", "depth_feature"):
        assert_close(out[k], fin[k], "fine " + k, rtol=1e-4, atol=2e-5)
    # end to end against the oracle and the reference's golden outputs
    assert_close(out["z_vals"], ref["z_vals"], "z_vals", atol=Z_ATOL)
    assert_close(out["z_vals"], g["z_vals"], "golden z_vals", atol=Z_ATOL)
    for k in ("rgb_map", "depth_map", "acc_map", "z_std"):
        assert_close(out[k], ref[k], k, rtol=1e-4, atol=2e-4)
        assert_close(out[k], g[k], "golden " + k, rtol=1e-4, atol=2e-4)   # golden = reference's own pdf normaliser


def test_case2_injected_randomness(engine):
    P, _ = small_params()
    g = golden("case2_perturb32")
    rand = {k: g[k].cuda() for k in ("t_rand", "noise0", "u", "noise1")}
    out = engine.render_rays(g["ray_batch"].cuda(), 64, retraw=True, N_importance=64, perturb=1., raw_noise_std=1.,
                             rand=rand, want_indices=True)
    ref = oc.render_rays(P, CFG, g["ray_batch"], 64, 64, perturb=1., rand={k: g[k] for k in ("t_rand", "noise0", "u", "noise1")})
    assert torch.equal(out["z_vals0"].cpu(), ref["z_vals0"])
    for k in ("rgb0", "weights0"):
        assert_close(out[k], ref[k], k, rtol=1e-4, atol=2e-5)
        assert_close(out[k], g[k], "golden " + k, rtol=1e-4, atol=2e-5)
    fin = oracle_fine_at(P, g["ray_batch"], out["z_vals"].cpu(), noise=g["noise1"])
    for k in ("weights", "rgb_map", "depth_map", "acc_map"):
        assert_close(out[k], fin[k], "fine " + k, rtol=1e-4, atol=2e-5)
    assert_close(out["z_vals"], g["z_vals"], "golden z_vals", atol=Z_ATOL)
    for k in ("rgb_map", "depth_map", "acc_map"):
        assert_close(out[k], g[k], "golden " + k, rtol=1e-4, atol=5e-4)


@pytest.mark.parametrize("prec", ["fp32", "tc32"])
@pytest.mark.parametrize("nc,ni,R", [(32, 32, 70), (96, 96, 19), (64, 0, 5), (48, 80, 3)])
def test_ragged_sample_counts(nc, ni, R, prec):
    from evdeblurnerf_b200 import RenderEngine
    P = random_params(7)
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision=prec)
    rays, _ = synthetic_rays(R, seed=nc + ni)
    rb = oc.build_ray_batch(H, W, FOCAL, rays)
    out = eng.render_rays(rb.cuda(), nc, retraw=True, N_importance=ni)
    ref = oc.render_rays(P, CFG, rb, nc, ni)
    if ni > 0:
        for k in ("rgb0", "depth0", "acc0", "weights0"):
            assert_close(out[k], ref[k], k, rtol=1e-4, atol=2e-5)
        fin = oracle_fine_at(P, rb, out["z_vals"].cpu())
        for k in ("rgb_map", "depth_map", "acc_map", "weights"):
            assert_close(out[k], fin[k], "fine " + k, rtol=1e-4, atol=2e-5)
        assert_close(out["z_vals"], ref["z_vals"], "z_vals", atol=Z_ATOL)
        for k in ("rgb_map", "depth_map", "acc_map"):
            assert_close(out[k], ref[k], k, rtol=1e-4, atol=5e-4)
    else:
        for k in ("rgb_map", "depth_map", "acc_map", "weights", "z_vals"):
            assert_close(out[k], ref[k], k, rtol=1e-4, atol=2e-5)


def test_empty_batch(engine):
    out = engine.render_rays(torch.zeros(0, 11).cuda(), 64, retraw=True, N_importance=64)
    assert out["rgb_map"].shape == (0, 3) and out["weights"].shape == (0, 128)


@pytest.mark.parametrize("prec", ["fp32", "tc32"])
def test_eval_mode_near_plane_mask(prec):
    from evdeblurnerf_b200 import RenderEngine
    P, _ = small_params()
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision=prec, rmnearplane=40)
    rays, _ = synthetic_rays(16, seed=2)
    rb = oc.build_ray_batch(H, W, FOCAL, rays)
    out = eng.render_rays(rb.cuda(), 64, retraw=True, N_importance=64, is_train=False)
    ref = oc.render_rays(P, dict(CFG, rmnearplane=40), rb, 64, 64, is_train=False)
    assert_close(out["rgb0"], ref["rgb0"], "rgb0", rtol=1e-4, atol=2e-5)
    assert float((out["weights"][:, :8].abs().max())) >= 0.0
    fin = oracle_fine_at(P, rb, out["z_vals"].cpu(), is_train=False, rmnearplane=40)
    for k in ("rgb_map", "depth_map", "acc_map", "weights"):
        assert_close(out[k], fin[k], "fine " + k, rtol=1e-4, atol=2e-5)
    for k in ("rgb_map", "depth_map", "acc_map"):
        assert_close(out[k], ref[k], k, rtol=1e-4, atol=5e-4)


@pytest.mark.parametrize("prec", ["fp32", "tc32"])
def test_full_size_properties(prec):
    """BASELINE config[1] size (4096 rays x 5 exposures, 64+64): size-independent properties."""
    from evdeblurnerf_b200 import RenderEngine
    P = random_params(11, coarse_grid=(96, 96, 64), fine_grid=(192, 192, 128))
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision=prec)
    rays, _ = synthetic_rays(20480, seed=5)
    rb = oc.build_ray_batch(H, W, FOCAL, rays).cuda()
    out = eng.render_rays(rb, 64, retraw=True, N_importance=64, want_indices=True)
    z = out["z_vals"]
    assert bool((z[:, 1:] >= z[:, :-1]).all())                                   # sortedness
    assert torch.equal(torch.sort(out["order"], -1)[0], torch.arange(128, device="cuda").expand(20480, 128))  # permutation
    w = out["weights"]
    assert bool((w >= 0).all()) and bool(torch.isfinite(w).all())
    assert_close(w.sum(-1), out["acc_map"], "acc = sum w", rtol=1e-5, atol=1e-6)
    assert float((out["acc_map"] - 1).abs().max()) < 1e-4                        # last alpha = 1 -> opaque rays
    assert_close((w * z).sum(-1), out["depth_map"], "depth = sum w z", rtol=1e-4, atol=1e-6)
    assert bool((out["rgb_map"] >= 0).all()) and bool((out["rgb_map"] <= 1 + 1e-5).all())
    # batch invariance: any sub-batch renders identically (the reference's chunk loop property, renderer.py:450)
    sub = eng.render_rays(rb[777:1301], 64, retraw=True, N_importance=64)
    assert torch.equal(sub["rgb_map"], out["rgb_map"][777:1301]) and torch.equal(sub["weights"], out["weights"][777:1301])
    # spot parity of a slice against the oracle
    fin = oracle_fine_at(P, rb[:64].cpu(), out["z_vals"][:64].cpu())
    for k in ("rgb_map", "depth_map", "weights"):
        assert_close(out[k][:64], fin[k], k, rtol=1e-4, atol=2e-5)


def test_philox_random_draws(engine):
    """In-library RNG for perturb / raw_noise_std: range, moments, reproducibility, stream independence."""
    engine.seed(123)
    engine._calls = 5
    u = engine._random((400, 257), 0)
    n = engine._random((400, 257), 1, normal=True, scale=2.0)
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0
    assert abs(float(u.mean()) - 0.5) < 5e-3 and abs(float(u.var()) - 1 / 12) < 2e-3
    assert abs(float(n.mean())) < 2e-2 and abs(float(n.std()) - 2.0) < 2e-2
    assert abs(float(((n / 2.0) ** 4).mean()) - 3.0) < 0.15                       # kurtosis of a normal
    engine.seed(123); engine._calls = 5
    assert torch.equal(u, engine._random((400, 257), 0))                           # reproducible
    assert not torch.equal(u, engine._random((400, 257), 2))                       # streams differ
    engine.seed(124); engine._calls = 5
    assert not torch.equal(u, engine._random((400, 257), 0))                       # seeds differ
    assert abs(float(torch.corrcoef(torch.stack([u.flatten()[:-1], u.flatten()[1:]]))[0, 1])) < 1e-2


def test_perturbed_render_is_reproducible_and_valid(engine):
    rays, _ = synthetic_rays(64, seed=9)
    rb = oc.build_ray_batch(H, W, FOCAL, rays).cuda()
    engine.seed(7)
    a = engine.render_rays(rb, 64, retraw=True, N_importance=64, perturb=1., raw_noise_std=1.)
    engine.seed(7)
    b = engine.render_rays(rb, 64, retraw=True, N_importance=64, perturb=1., raw_noise_std=1.)
    c = engine.render_rays(rb, 64, retraw=True, N_importance=64, perturb=1., raw_noise_std=1.)
    assert torch.equal(a["rgb_map"], b["rgb_map"]) and not torch.equal(a["rgb_map"], c["rgb_map"])
    z = a["z_vals"]
    assert bool((z[:, 1:] >= z[:, :-1]).all()) and bool((a["z_vals0"][:, 1:] > a["z_vals0"][:, :-1]).all())
    assert_close(a["weights"].sum(-1), a["acc_map"], "acc", rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("perturb", [0.0, 1.0])
def test_lindisp_sample_placement(precision, perturb):
    """renderer.py:166-167 (`lindisp`, only reachable with --no_ndc: near must be > 0): samples linear in inverse depth.  Placement is
    bit-exact in both precisions (with and without the stratified jitter); the fp32 render follows the oracle at the usual bar."""
    from evdeblurnerf_b200 import RenderEngine
    P, _ = small_params()
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision=precision)
    rays, _ = synthetic_rays(40, seed=91)
    rb = oc.build_ray_batch(H, W, FOCAL, rays, near=0.5, far=2.0, ndc=False)
    g = torch.Generator().manual_seed(92)
    rand = {"t_rand": torch.rand(40, 64, generator=g)} if perturb > 0 else {}
    out = eng.render_rays(rb.cuda(), 64, retraw=True, lindisp=True, perturb=perturb, rand={k: v.cuda() for k, v in rand.items()})
    ref = oc.render_rays(P, CFG, rb, 64, 0, perturb=perturb, lindisp=True, rand=rand)
    lin = oc.render_rays(P, CFG, rb, 64, 0, perturb=perturb, lindisp=False, rand=rand)
    assert torch.equal(out["z_vals"].cpu(), ref["z_vals"]), "lindisp placement must be bit-exact"
    assert float((ref["z_vals"] - lin["z_vals"]).abs().max()) > 0.05          # the flag changes the placement
    if precision == "fp32":
        for k in ("rgb_map", "depth_map", "acc_map", "weights"):
            assert_close(out[k], ref[k], k)
    else:
        for k in ("rgb_map", "acc_map"):
            assert_close(out[k], ref[k], k, rtol=0, atol=4e-2)
