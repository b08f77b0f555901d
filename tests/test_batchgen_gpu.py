"""GPU parity of the device-side batch generation against the golden vectors written from the unmodified
reference's own functions (oracle/make_golden_batchgen.py) and the oracle.  Rays bit-exact (no FMA contraction), successor
walk exact (integers), interpolated poses within 2e-6 (float64 SLERP / spline against scipy)."""
import numpy as np
import pytest
import torch

import evdeblur_oracle as oc
from util import assert_close, golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    return golden("case7_batchgen")


def test_rgb_batch_matches_reference_getitem(g):
    from evdeblurnerf_b200.batchgen import RayBatchSampler
    K = g["K"].tolist()
    ds = RayBatchSampler(g["images"], g["poses"], K)
    assert len(ds) == g["images"].shape[0] * g["images"].shape[1] * g["images"].shape[2]
    b = ds[g["ray_ids"]]
    assert torch.equal(b["rays"].cpu(), g["rays"])                    # bit-exact ray origins / directions
    assert torch.equal(b["rgbsf"].cpu(), g["rgbsf"])
    assert torch.equal(b["images_idx"].cpu(), g["images_idx"])
    assert torch.equal(b["rays_x"].cpu(), g["rays_x"].float()) and torch.equal(b["rays_y"].cpu(), g["rays_y"].float())
    assert torch.equal(b["poses"].cpu(), g["poses"][g["images_idx"].reshape(-1)])
    assert ds[[]]["rays"].shape == (0, 3, 2)


def test_get_rays_pix_broadcast_pose_and_halfpix(g):
    from evdeblurnerf_b200.batchgen import get_rays_pix
    K = g["K"].tolist()
    gen = torch.Generator().manual_seed(1)
    xy = torch.rand(33, 2, generator=gen) * 15
    for half in (True, False):
        ref = oc.rays_from_pixels(xy, K, g["poses"][2], add_halfpix=half)
        o, d = get_rays_pix(xy.cuda(), K, g["poses"][2].cuda(), add_halfpix=half)
        assert_close(torch.stack([o, d], -1), ref, f"broadcast pose, halfpix={half}", rtol=1e-6, atol=1e-7)


def test_gather_successor_exact(g):
    from evdeblurnerf_b200.batchgen import gather_successor
    idx, neg, pos = gather_successor(g["q_idx"].cuda(), g["q_hops"].cuda(), g["succ"].cuda(), g["pol"].cuda())
    assert torch.equal(idx.cpu(), g["succ_idx"]) and torch.equal(neg.cpu(), g["neg"]) and torch.equal(pos.cpu(), g["pos"])
    # larger random graph against the oracle
    gen = torch.Generator().manual_seed(3)
    n_ev = 5000
    succ = torch.arange(n_ev) + torch.randint(1, 40, (n_ev,), generator=gen)
    succ[succ >= n_ev] = -1
    pol = (torch.randint(0, 2, (n_ev,), generator=gen) * 2 - 1).int()
    q, hops = torch.randint(0, n_ev, (700,), generator=gen), torch.randint(0, 12, (700,), generator=gen)
    r = oc.gather_successor(q, hops, succ, pol)
    o = gather_successor(q.cuda(), hops.cuda(), succ.cuda(), pol.cuda())
    for a, b in zip(o, r):
        assert torch.equal(a.cpu(), b)


def test_pose_interpolation_matches_scipy_pipeline(g):
    from evdeblurnerf_b200.batchgen import PoseInterpolator, get_rays_pix
    interp = PoseInterpolator(g["times"].numpy(), g["rots"].numpy(), g["trans"].numpy(), bd_scale=float(g["bd_scale"]),
                              recenter_c2w=g["recenter_c2w"].numpy())
    poses = interp(g["tq"].numpy())
    assert_close(poses, g["event_poses"], "interpolated event poses", rtol=0, atol=2e-6)
    # no recentering, unit scale, against the oracle directly
    plain = PoseInterpolator(g["times"].numpy(), g["rots"].numpy(), g["trans"].numpy())
    ref = oc.interpolate_event_poses(oc.pose_interpolator(g["times"].numpy(), g["rots"].numpy(), g["trans"].numpy()), g["tq"].numpy())
    assert_close(plain(g["tq"].numpy()), ref, "plain poses", rtol=0, atol=2e-6)
    # the event rays the reference builds from them (loader_events.py:291-296)
    o, d = get_rays_pix(g["ev_xy"].cuda(), g["K"].tolist(), torch.as_tensor(g["event_poses"]).cuda())
    assert torch.equal(torch.stack([o, d], -1).cpu(), g["ev_rays"])


def test_event_batch_sampler_against_oracle_composition(g):
    """sample_events end to end: successor walk (or the plain successor), pose interpolation at the start / end timestamps,
    rays at the event pixel -- both accumulation modes."""
    from evdeblurnerf_b200.batchgen import EventBatchSampler, PoseInterpolator
    gen = torch.Generator().manual_seed(5)
    times, rots, trans = g["times"].numpy(), g["rots"].numpy(), g["trans"].numpy()
    n_ev, n_ids = 400, 50
    ts = np.sort(np.random.default_rng(2).uniform(times[0], times[-1], n_ev))
    coord_id = torch.randint(0, n_ids, (n_ev,), generator=gen)
    succ = torch.arange(n_ev) + torch.randint(1, 6, (n_ev,), generator=gen)
    succ[succ >= n_ev] = 0                                            # keep every successor valid for this test
    pol = torch.randint(0, 2, (n_ev,), generator=gen) * 2 - 1
    events = torch.stack([coord_id.double(), torch.zeros(n_ev).double(), torch.as_tensor(ts), pol.double(), succ.double()], -1)
    id_to_coords = torch.randint(0, 14, (n_ids, 2), generator=gen).float()
    K = g["K"].tolist()
    interp = PoseInterpolator(times, rots, trans, bd_scale=0.5)
    o_interp = oc.pose_interpolator(times, rots, trans)
    ids = torch.randint(0, n_ev, (37,), generator=gen)
    for accum in ((0, 0), (1, 4)):
        ds = EventBatchSampler(events, id_to_coords, K, interp, num_successors=torch.full((n_ev,), 10), accum_steps=lambda s, a=accum: a)
        hops = torch.randint(0, 4, (37,), generator=gen) if accum != (0, 0) else None
        out = ds.sample_events(ids, 0, sampled_hops=None if hops is None else hops.cuda())
        if hops is None:
            end = succ[ids]
            p = pol[end]
            pos, neg = torch.where(p > 0, p, 0), torch.where(p > 0, 0, p)
        else:
            end, neg, pos = oc.gather_successor(ids, hops, succ, pol.int())
        assert torch.equal(out["events_pos_pol_cumsum"].cpu().long(), pos.long()) and torch.equal(out["events_neg_pol_cumsum"].cpu().long(), neg.long())
        for key, tq in (("events_rays_start", ts[ids.numpy()]), ("events_rays_end", ts[end.numpy()])):
            poses = torch.as_tensor(oc.interpolate_event_poses(o_interp, tq, 0.5))
            ref = oc.rays_from_pixels(id_to_coords[coord_id[ids]], K, poses)
            assert_close(out[key], ref, key, rtol=1e-5, atol=5e-6)
        assert torch.equal(out["events_coords_ids"].cpu(), coord_id[ids])
