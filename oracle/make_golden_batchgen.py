"""TEST INFRASTRUCTURE ONLY (this container only: imports the unmodified reference from /root/reference).

Pins the batch-generation restatements of oracle/evdeblur_oracle.py (get_rays_pix, LLFFDataset.__getitem__ arithmetic,
gather_successor, the SLERP + cubic pose interpolator + recenter) against the reference's own functions on seeded inputs
and writes tests/golden/case7_batchgen.npz.

    python oracle/make_golden_batchgen.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import evdeblur_oracle as oc  # noqa: E402
from reference_harness import _install_shims  # noqa: E402


def main():
    _install_shims()
    from utils.rays import get_rays_pix
    from utils.events import gather_successor
    from utils.data import _get_slerp_interpolator, recenter_poses
    from utils.misc import unravel_index
    from scipy.spatial.transform import Rotation
    g = torch.Generator().manual_seed(0)
    rng = np.random.default_rng(0)
    out = {}
    # ---- camera poses on a smooth trajectory ---------------------------------------------------------------------------
    Kn = 12
    times = np.cumsum(rng.uniform(0.8e5, 1.2e5, Kn)) + 1.0e9                     # us, float64
    rots = Rotation.from_rotvec(np.cumsum(rng.normal(0, 0.08, (Kn, 3)), 0)).as_matrix()
    trans = np.cumsum(rng.normal(0, 0.05, (Kn, 3)), 0)
    n_img, Hh, Ww = 5, 14, 18
    K = [[20.0, 0, 9.0], [0, 21.0, 7.0], [0, 0, 1.0]]
    poses = torch.tensor(np.concatenate([rots[:n_img], trans[:n_img, :, None]], -1), dtype=torch.float32)
    images = torch.rand(n_img, Hh, Ww, 3, generator=g)
    # ---- RGB batch: unravel + get_rays_pix + gathers --------------------------------------------------------------------------
    ray_ids = torch.randint(0, n_img * Hh * Ww, (40,), generator=g)
    img_id, ry, rx = unravel_index(ray_ids, (n_img, Hh, Ww)).T
    ro, rd = get_rays_pix(torch.stack([rx, ry], -1), K, poses[img_id])
    ref_rays = torch.stack([ro, rd], dim=-2).permute(0, 2, 1)
    b = oc.make_rgb_batch(ray_ids, images, poses, K)
    assert torch.equal(b["images_idx"].reshape(-1), img_id) and torch.allclose(b["rays"], ref_rays, rtol=0, atol=0)
    assert torch.equal(b["rgbsf"], images[img_id, ry, rx])
    out.update(ray_ids=ray_ids, images=images, poses=poses, K=torch.tensor(K), rays=ref_rays, rgbsf=b["rgbsf"],
               rays_x=b["rays_x"], rays_y=b["rays_y"], images_idx=b["images_idx"])
    # ---- successor walk ------------------------------------------------------------------------------------------------------
    n_ev = 300
    succ = torch.arange(n_ev) + torch.randint(1, 9, (n_ev,), generator=g)
    succ[succ >= n_ev] = -1                                                        # chain ends
    pol = torch.randint(0, 2, (n_ev,), generator=g).int() * 2 - 1
    q_idx = torch.randint(0, n_ev, (64,), generator=g)
    q_hops = torch.randint(0, 6, (64,), generator=g)
    r_idx, r_neg, r_pos = gather_successor(q_idx, q_hops, succ, pol)
    o_idx, o_neg, o_pos = oc.gather_successor(q_idx, q_hops, succ, pol)
    assert torch.equal(r_idx, o_idx) and torch.equal(r_neg, o_neg) and torch.equal(r_pos, o_pos)
    assert int((r_idx < 0).sum()) > 0 and int((r_idx >= 0).sum()) > 0
    out.update(succ=succ, pol=pol, q_idx=q_idx, q_hops=q_hops, succ_idx=r_idx, neg=r_neg, pos=r_pos)
    # ---- pose interpolation (data/loader_events.py:133-148, 175-183; utils/data.py:34-61, 167-183) ----------------------------
    interp_ref = _get_slerp_interpolator(times, rots, trans)
    tq = np.concatenate([rng.uniform(times[0] - 1e4, times[-1] + 1e4, 50), times[[0, 3, -1]]])
    tq_c = np.clip(tq, times.min(), times.max())
    irots, itrans = interp_ref(tq_c)
    bottom = np.array([0, 0, 0, 1]).reshape(1, 1, -1).repeat(tq.shape[0], axis=0)
    int_poses = np.block([[irots, itrans[..., np.newaxis]], [bottom]])
    int_poses = np.concatenate([int_poses[..., 1:2], -int_poses[..., 0:1], int_poses[..., 2:]], -1).astype(np.float32)
    bd_scale = 0.37
    int_poses[..., :3, 3] *= bd_scale
    c2w = np.eye(4)
    c2w[:3, :3] = Rotation.from_rotvec([0.1, -0.2, 0.05]).as_matrix()
    c2w[:3, 3] = [0.3, -0.1, 0.2]
    ref_poses = recenter_poses(int_poses, c2w=c2w)[:, :3, :4]
    mine = oc.interpolate_event_poses(oc.pose_interpolator(times, rots, trans), tq, bd_scale, c2w)
    assert np.allclose(mine, ref_poses, rtol=0, atol=1e-6), float(np.abs(mine - ref_poses).max())
    out.update(times=times, rots=rots, trans=trans, tq=tq, bd_scale=np.float64(bd_scale), recenter_c2w=c2w, event_poses=ref_poses)
    # event rays at the interpolated poses (loader_events.py:291-296)
    ev_xy = torch.tensor(rng.integers(0, [Ww, Hh], (tq.shape[0], 2)), dtype=torch.float32)
    eo, ed = get_rays_pix(ev_xy, K, torch.tensor(ref_poses), add_halfpix=True)
    out.update(ev_xy=ev_xy, ev_rays=torch.stack([eo, ed], 1).permute(0, 2, 1))
    path = os.path.join(HERE, "..", "tests", "golden", "case7_batchgen.npz")
    np.savez_compressed(path, **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()})
    print("wrote", os.path.normpath(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
