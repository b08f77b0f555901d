"""GPU parity tests of the tcgen05 (bf16 operands, fp32 accumulate) fine pass.

Two checks: (1) tight, against a torch fp32 reference that applies the same bf16 operand roundings
(util.emulated_bf16_fine) -- differences are accumulation order only; (2) loose, against the fp32 oracle, with the
tolerance bf16 operands allow (this mode is judged by image quality, SURVEY.md section 7 "bf16 vs the 1e-4 bar")."""
import pytest
import torch

import evdeblur_oracle as oc
from util import (AABB, CFG, FOCAL, H, W, assert_close, emulated_bf16_coarse, emulated_bf16_fine, oracle_fine_at,
                  random_params, small_params, synthetic_rays)

pytestmark = pytest.mark.gpu

# tolerances, stated: (1) emulated reference: rgb/weights 2e-3 abs (bf16 rounding flips of intermediate activations);
# (2) fp32 oracle: 3e-2 abs on colours / weights (bf16 has 8 mantissa bits; K up to 256 per layer)
EMU_ATOL, ORACLE_ATOL = 3e-3, 4e-2


@pytest.fixture(scope="module")
def engines():
    from evdeblurnerf_b200 import RenderEngine
    P, _ = small_params()
    Pc = {k: v.cuda() for k, v in P.items()}
    return P, RenderEngine(Pc, *AABB, precision="bf16"), RenderEngine(Pc, *AABB, precision="fp32")


def test_tc_coarse_matches_emulated_reference(engines):
    P, eng, _ = engines
    rays, _ = synthetic_rays(150, seed=33)
    rb = oc.build_ray_batch(H, W, FOCAL, rays)
    out = eng.render_rays(rb.cuda(), 64, retraw=True, use_awp=True)          # coarse only, feature_map requested -> full schedule
    emu = emulated_bf16_coarse(P, rb, 64, lean=False)
    assert torch.equal(out["z_vals"].cpu(), emu["z_vals"])                    # placement stays bit-exact
    assert_close(out["depth_feature"], emu["feature"], "geo", rtol=2e-2, atol=5e-3)
    for k in ("weights", "rgb_map", "depth_map", "acc_map"):
        assert_close(out[k], emu[k], k, rtol=0, atol=EMU_ATOL)
    ref = oc.render_rays(P, CFG, rb, 64, 0)
    for k in ("weights", "rgb_map", "depth_map", "acc_map"):
        assert_close(out[k], ref[k], "fp32 " + k, rtol=0, atol=ORACLE_ATOL)
    lean = eng.render_rays(rb.cuda(), 64, retraw=True)                          # default: lean (folded) schedule
    emu_l = emulated_bf16_coarse(P, rb, 64, lean=True)
    for k in ("weights", "rgb_map", "depth_map", "acc_map"):
        assert_close(lean[k], emu_l[k], "lean " + k, rtol=0, atol=EMU_ATOL)
        assert_close(lean[k], ref[k], "lean fp32 " + k, rtol=0, atol=ORACLE_ATOL)


@pytest.mark.parametrize("nc,R", [(32, 41), (64, 7), (96, 10), (128, 5), (48, 9)])
def test_tc_coarse_ragged(nc, R):
    from evdeblurnerf_b200 import RenderEngine
    P = random_params(13)
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision="bf16")
    rays, _ = synthetic_rays(R, seed=nc)
    rb = oc.build_ray_batch(H, W, FOCAL, rays)
    g = torch.Generator().manual_seed(nc)
    t_rand, noise = torch.rand(R, nc, generator=g), torch.randn(R, nc - 1, generator=g)
    out = eng.render_rays(rb.cuda(), nc, retraw=True, perturb=1., raw_noise_std=1., rand={"t_rand": t_rand.cuda(), "noise0": noise.cuda()})
    emu = emulated_bf16_coarse(P, rb, nc, t_rand=t_rand, noise=noise)
    assert torch.equal(out["z_vals"].cpu(), emu["z_vals"])
    for k in ("weights", "rgb_map", "depth_map", "acc_map"):
        assert_close(out[k], emu[k], k, rtol=0, atol=EMU_ATOL)


def test_tc_fine_matches_emulated_reference(engines):
    P, eng, _ = engines
    rays, _ = synthetic_rays(150, seed=31)
    rb = oc.build_ray_batch(H, W, FOCAL, rays)
    out = eng.render_rays(rb.cuda(), 64, retraw=True, N_importance=64, use_awp=True)
    z = out["z_vals"].cpu()
    emu = emulated_bf16_fine(P, rb, z)
    assert_close(out["depth_feature"], emu["depth_feature"], "geo", rtol=2e-2, atol=5e-3)
    for k in ("weights", "rgb_map", "depth_map", "acc_map"):
        assert_close(out[k], emu[k], k, rtol=0, atol=EMU_ATOL)
    fin = oracle_fine_at(P, rb, z)
    for k in ("weights", "rgb_map", "depth_map", "acc_map"):
        assert_close(out[k], fin[k], "fp32 " + k, rtol=0, atol=ORACLE_ATOL)


def test_tc_fine_lean_schedule_matches_emulated_reference(engines):
    """Default (no depth_feature) path: basis_mat / sigma_net.1 folded into the neighbouring layers at pack time."""
    P, eng, _ = engines
    rays, _ = synthetic_rays(150, seed=34)
    rb = oc.build_ray_batch(H, W, FOCAL, rays)
    out = eng.render_rays(rb.cuda(), 64, retraw=True, N_importance=64)
    z = out["z_vals"].cpu()
    emu = emulated_bf16_fine(P, rb, z, lean=True)
    for k in ("weights", "rgb_map", "depth_map", "acc_map"):
        assert_close(out[k], emu[k], k, rtol=0, atol=EMU_ATOL)
    fin = oracle_fine_at(P, rb, z)
    for k in ("weights", "rgb_map", "depth_map", "acc_map"):
        assert_close(out[k], fin[k], "fp32 " + k, rtol=0, atol=ORACLE_ATOL)


def test_tc_noise_and_eval_mask(engines):
    P, _, _ = engines
    from evdeblurnerf_b200 import RenderEngine
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision="bf16", rmnearplane=40)
    rays, _ = synthetic_rays(40, seed=32)
    rb = oc.build_ray_batch(H, W, FOCAL, rays)
    g = torch.Generator().manual_seed(1)
    rand = {"t_rand": torch.rand(40, 64, generator=g), "u": torch.rand(40, 64, generator=g),
            "noise0": torch.randn(40, 63, generator=g), "noise1": torch.randn(40, 127, generator=g)}
    out = eng.render_rays(rb.cuda(), 64, retraw=True, N_importance=64, perturb=1., raw_noise_std=1.,
                          rand={k: v.cuda() for k, v in rand.items()}, is_train=False)
    emu = emulated_bf16_fine(P, rb, out["z_vals"].cpu(), noise=rand["noise1"], is_train=False, rmnearplane=40, lean=True)
    for k in ("weights", "rgb_map", "depth_map", "acc_map"):
        assert_close(out[k], emu[k], k, rtol=0, atol=EMU_ATOL)


@pytest.mark.parametrize("nc,ni,R", [(32, 32, 37), (64, 17, 9), (64, 64, 1), (96, 96, 21), (128, 128, 5), (64, 70, 4)])
def test_tc_ragged_sample_counts(nc, ni, R):
    from evdeblurnerf_b200 import RenderEngine
    P = random_params(9)
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision="bf16")
    rays, _ = synthetic_rays(R, seed=nc + ni)
    rb = oc.build_ray_batch(H, W, FOCAL, rays)
    out = eng.render_rays(rb.cuda(), nc, retraw=True, N_importance=ni)
    emu = emulated_bf16_fine(P, rb, out["z_vals"].cpu(), lean=True)
    for k in ("weights", "rgb_map", "depth_map", "acc_map"):
        assert_close(out[k], emu[k], k, rtol=0, atol=EMU_ATOL)


def test_tc_full_size_properties():
    from evdeblurnerf_b200 import RenderEngine
    P = random_params(11, coarse_grid=(96, 96, 64), fine_grid=(192, 192, 128))
    Pc = {k: v.cuda() for k, v in P.items()}
    eng, eng32 = RenderEngine(Pc, *AABB, precision="bf16"), RenderEngine(Pc, *AABB, precision="fp32")
    rays, _ = synthetic_rays(20480, seed=5)
    rb = oc.build_ray_batch(H, W, FOCAL, rays).cuda()
    out = eng.render_rays(rb, 64, retraw=True, N_importance=64)
    w, z = out["weights"], out["z_vals"]
    assert bool(torch.isfinite(w).all()) and bool((w >= 0).all())
    assert_close(w.sum(-1), out["acc_map"], "acc = sum w", rtol=1e-5, atol=1e-5)
    assert float((out["acc_map"] - 1).abs().max()) < 1e-4
    assert_close((w * z).sum(-1), out["depth_map"], "depth = sum w z", rtol=1e-4, atol=1e-5)
    sub = eng.render_rays(rb[4000:4300], 64, retraw=True, N_importance=64)
    assert torch.equal(sub["rgb_map"], out["rgb_map"][4000:4300])          # batch / CTA-assignment invariance
    ref = eng32.render_rays(rb, 64, retraw=True, N_importance=64)
    # the bf16 coarse pass shares the fp32 coarse kernel only through bf16 planes, so compare image-level statistics
    err = (out["rgb_map"] - ref["rgb_map"]).abs()
    assert float(err.mean()) < 1e-2 and float(err.max()) < 0.15
    emu = emulated_bf16_fine(P, rb[:48].cpu(), z[:48].cpu(), lean=True)
    for k in ("weights", "rgb_map"):
        assert_close(out[k][:48], emu[k], k, rtol=0, atol=EMU_ATOL)


@pytest.mark.parametrize("precision", ["bf16", "tc32"])
def test_tc_coarse_slow_producer_does_not_stall_the_pipeline(precision):
    """Regression (found by the 2-GPU bench run): with perturbed sample placement the producer warps of the decoupled coarse kernels
    are slower than the MLP chain, the epilogue then finishes tile k - 1 before the producer asks about tile k - 2, and a 1-bit mbarrier
    parity cannot tell "two phases ahead" from "not completed" -- the kernel trapped in its bounded wait.  Many tiles per CTA, perturb
    and noise on, repeated with fresh draws; the results stay finite and composited weights sum to the accumulated opacity."""
    from evdeblurnerf_b200 import RenderEngine
    P = random_params(13, coarse_grid=(96, 96, 64), fine_grid=(192, 192, 128))
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision=precision)
    for seed in (21, 22):
        rays, _ = synthetic_rays(20480, seed=seed)
        rb = oc.build_ray_batch(H, W, FOCAL, rays).cuda()
        for _ in range(6):
            out = eng.render_rays(rb, 64, retraw=True, N_importance=0, perturb=1., raw_noise_std=1.)
            torch.cuda.synchronize()
        w = out["weights"]
        assert bool(torch.isfinite(w).all()) and bool(torch.isfinite(out["rgb_map"]).all())
        assert_close(w.sum(-1), out["acc_map"], "acc = sum w", rtol=1e-5, atol=1e-5)
