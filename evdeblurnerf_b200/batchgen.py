"""Host mirror of the reference's batch generation: `LLFFDataset.__getitem__` (data/loader.py:325-356),
`get_rays_pix` (utils/rays.py:25-36), `LLFFEventsDataset.sample_events` / `interpolate_poses` (data/loader_events.py:133-148,
259-304), `gather_successor` (utils/events.py:221-257) -- on the device, so that a renderer that is an order of magnitude
faster is not starved by DataLoader workers and a per-step numpy / scipy pose interpolation on the host.

The scene tensors (images, poses, the event stream with its successor map) live on the GPU; every call launches one small
kernel per product.  scipy is used ONCE, at construction, to turn the translation spline into piecewise-cubic coefficients.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

HALF_PIX = 0.5      # utils/rays.py:5


def get_rays_pix(coords, K, c2ws, add_halfpix=True):
    """utils/rays.py:25-36: coords [n,2] = (x, y); c2ws [n,3,4] or [3,4] -> (rays_o [n,3], rays_d [n,3])."""
    xy = coords.detach().to(torch.float32).contiguous()
    c2 = c2ws.detach().to(torch.float32)[..., :3, :4].contiguous()
    n = xy.shape[0]
    rays = torch.empty((n, 3, 2), dtype=torch.float32, device=xy.device)
    check(_lib.load().edn_rays_from_pixels(ptr(xy), ptr(c2), 1 if c2.ndim == 2 else 0, n, float(K[0][0]), float(K[1][1]), float(K[0][2]),
                                           float(K[1][2]), 1 if add_halfpix else 0, ptr(rays), stream_ptr()), "edn_rays_from_pixels")
    return rays[..., 0], rays[..., 1]


class RayBatchSampler:
    """`LLFFDataset.__getitem__` over device-resident images [n_img,H,W,3] and poses [n_img,3,4]."""

    def __init__(self, images, poses, K, device="cuda"):
        self.images = torch.as_tensor(images, dtype=torch.float32).to(device).contiguous()
        self.poses = torch.as_tensor(poses, dtype=torch.float32)[:, :3, :4].to(device).contiguous()
        self.n_imgs, self.h, self.w = self.images.shape[:3]
        self.K = [[float(v) for v in row] for row in K]
        self.n_rays = self.n_imgs * self.h * self.w

    def __len__(self):
        return self.n_rays

    def __getitem__(self, ray_ids):
        ids = torch.as_tensor(ray_ids, dtype=torch.int64, device=self.images.device).reshape(-1).contiguous()
        n, dev = ids.shape[0], ids.device
        f32 = dict(dtype=torch.float32, device=dev)
        out = {"rays": torch.empty((n, 3, 2), **f32), "rays_x": torch.empty((n, 1), **f32), "rays_y": torch.empty((n, 1), **f32),
               "images_idx": torch.empty((n, 1), dtype=torch.int64, device=dev), "rgbsf": torch.empty((n, 3), **f32),
               "poses": torch.empty((n, 3, 4), **f32)}
        K = self.K
        check(_lib.load().edn_make_rgb_batch(ptr(ids), n, ptr(self.images), ptr(self.poses), self.n_imgs, self.h, self.w, K[0][0], K[1][1],
                                             K[0][2], K[1][2], ptr(out["rays"]), ptr(out["rays_x"]), ptr(out["rays_y"]), ptr(out["images_idx"]),
                                             ptr(out["rgbsf"]), ptr(out["poses"]), stream_ptr()), "edn_make_rgb_batch")
        return out


def gather_successor(query_idx, query_hops, successor_map, polarities):
    """utils/events.py:221-257 -> (successor idx, negative polarity sum, positive polarity sum)."""
    q = query_idx.to(torch.int64).contiguous()
    hops = query_hops.to(torch.int64).contiguous()
    succ = successor_map.to(torch.int64).contiguous()
    pol = polarities.to(torch.int32).contiguous()
    n = q.shape[0]
    out = torch.empty((n,), dtype=torch.int64, device=q.device)
    neg = torch.empty((n,), dtype=torch.int32, device=q.device)
    pos = torch.empty((n,), dtype=torch.int32, device=q.device)
    check(_lib.load().edn_gather_successor(ptr(q), ptr(hops), n, ptr(succ), ptr(pol), succ.shape[0], ptr(out), ptr(neg), ptr(pos), stream_ptr()),
          "edn_gather_successor")
    return out, neg, pos


class PoseInterpolator:
    """`events_pose_bspl` + `interpolate_poses` (data/loader_events.py:133-148, 175-183): SLERP of the rotations and the cubic
    translation spline of utils/data.py:34-61, the matrix-format change, `bd_scale` and `recenter_poses` -- evaluated on the
    device.  spherify is not built (no shipped event config uses it)."""

    def __init__(self, times_us, rots, trans, bd_scale=1.0, recenter_c2w=None, device="cuda"):
        from scipy.interpolate import PPoly, make_interp_spline
        from scipy.spatial.transform import Rotation
        t = np.asarray(times_us, dtype=np.float64)
        q = Rotation.from_matrix(np.asarray(rots, dtype=np.float64)).as_quat()           # (x, y, z, w)
        tr = np.asarray(trans, dtype=np.float64)
        pps = [PPoly.from_spline(make_interp_spline(t, tr[:, d], k=3)) for d in range(3)]   # == interp1d(kind="cubic"), per axis
        keep = np.diff(pps[0].x) > 0                                                      # drop the repeated boundary knots
        brk = np.concatenate([pps[0].x[:-1][keep], pps[0].x[-1:]])
        coef = np.ascontiguousarray(np.stack([pp.c[:, keep].T for pp in pps], -1))        # [interval][power][dim]
        dd = dict(dtype=torch.float64, device=device)
        self.times, self.quats = torch.tensor(t, **dd), torch.tensor(np.ascontiguousarray(q), **dd)
        self.brk, self.coef = torch.tensor(brk, **dd), torch.tensor(coef, **dd)
        self.bd_scale = float(bd_scale)
        self.recenter_inv = None if recenter_c2w is None else torch.tensor(np.linalg.inv(np.asarray(recenter_c2w, dtype=np.float64)), **dd)

    def __call__(self, t):
        tq = torch.as_tensor(t, dtype=torch.float64, device=self.times.device).reshape(-1).contiguous()
        out = torch.empty((tq.shape[0], 3, 4), dtype=torch.float32, device=tq.device)
        check(_lib.load().edn_interpolate_poses(ptr(tq), tq.shape[0], ptr(self.times), ptr(self.quats), self.times.shape[0], ptr(self.brk),
                                                ptr(self.coef), self.brk.shape[0], self.bd_scale, ptr(self.recenter_inv), ptr(out), stream_ptr()),
              "edn_interpolate_poses")
        return out


class EventBatchSampler:
    """`LLFFEventsDataset.sample_events` (data/loader_events.py:259-304) over a device-resident event stream.
    events [n_ev, 5] = (coord id, ..., timestamp us, polarity, successor idx) as the reference stores it (columns -3, -2, -1
    are read); id_to_coords [n_ids, 2]; the accumulation schedule is given by `accum_steps(global_step) -> (min, max)`."""

    def __init__(self, events, id_to_coords, K, pose_interpolator, num_successors=None, id_to_color_map=None, integer_coords=True,
                 accum_steps=lambda step: (0, 0), device="cuda"):
        ev = torch.as_tensor(events, dtype=torch.float64).to(device)
        self.coord_ids = ev[:, 0].to(torch.int64).contiguous()
        self.timestamps = ev[:, -3].contiguous()
        self.polarity = ev[:, -2].to(torch.int32).contiguous()
        self.successor = ev[:, -1].to(torch.int64).contiguous()
        self.num_successors = None if num_successors is None else torch.as_tensor(num_successors).to(device)
        self.id_to_coords = torch.as_tensor(id_to_coords, dtype=torch.float32).to(device).contiguous()
        self.id_to_color_map = None if id_to_color_map is None else torch.as_tensor(id_to_color_map).to(device)
        self.K, self.interp, self.integer_coords, self.accum_steps = K, pose_interpolator, integer_coords, accum_steps

    def sample_events(self, events_ids, global_step=0, sampled_hops=None):
        ids = torch.as_tensor(events_ids, dtype=torch.int64, device=self.successor.device).reshape(-1)
        min_step, max_step = (int(v) for v in self.accum_steps(global_step))
        if (min_step, max_step) != (0, 0):
            if sampled_hops is None:      # torch_randint_vec (utils/misc.py:87-92): uniform draw, rounded
                lo = torch.full_like(ids, min_step - 1, dtype=torch.float32)
                hi = torch.minimum(torch.tensor(float(max_step), device=ids.device), self.num_successors[ids].float()) - 1 + 1e-5
                sampled_hops = torch.round(lo + (hi - lo) * torch.rand(ids.shape[0], device=ids.device)).to(torch.int64)
            end_idx, neg, pos = gather_successor(ids, sampled_hops, self.successor, self.polarity)
        else:
            end_idx = self.successor[ids]
            p = self.polarity[end_idx]
            pos, neg = torch.where(p > 0, p, torch.zeros_like(p)), torch.where(p > 0, torch.zeros_like(p), p)
        poses_start = self.interp(self.timestamps[ids])
        poses_end = self.interp(self.timestamps[end_idx])
        coords_ids = self.coord_ids[ids]
        coords = self.id_to_coords[coords_ids]
        rs = torch.stack(get_rays_pix(coords, self.K, poses_start, add_halfpix=self.integer_coords), -1)
        re = torch.stack(get_rays_pix(coords, self.K, poses_end, add_halfpix=self.integer_coords), -1)
        return {"events_pos_pol_cumsum": pos, "events_neg_pol_cumsum": neg, "events_rays_start": rs, "events_rays_end": re,
                "events_coords_ids": coords_ids,
                "events_color_map": self.id_to_color_map[coords_ids] if self.id_to_color_map is not None else None}
