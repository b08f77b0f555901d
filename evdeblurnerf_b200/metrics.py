"""Eval-path image metrics: host mirror of utils/metrics.py::compute_img_metric for "mse", "psnr" and "ssim" on the
device the images live on (the reference moves them to the CPU and calls skimage).  Conventions follow the reference: images in
[0, 1] are mapped to [-1, 1] first (utils/metrics.py:44-46), so the data range is 2 -- PSNR = 10 log10(4 / mse) --, SSIM is skimage's
default estimator: 7 x 7 uniform windows, K1 = 0.01, K2 = 0.03, sample covariance, mean over the valid interior and the channels.
mse / psnr are pinned by tests/test_metrics_gpu.py against closed forms; SSIM restates skimage.metrics.structural_similarity
(v0.18, `multichannel=True`; skimage is not installed in this image, so that one is UNPINNED).  LPIPS needs the pretrained VGG
weights (no network here) and stays with the reference.  Not a hot path: torch tensor ops + edn_img2mse."""
import math

import torch

from .losses import img2mse


def _prep(im1, im2, fmt):
    a, b = (torch.as_tensor(x, dtype=torch.float32) for x in (im1, im2))
    a, b = (a * 2 - 1).clamp(-1, 1), (b * 2 - 1).clamp(-1, 1)
    if (a.dim() == 3 and fmt is None) or fmt in ("HWC", "CHW"):
        a, b = a[None], b[None]
    if (a.shape[-1] == 3 and fmt is None) or fmt in ("BHWC", "HWC"):
        a, b = a.permute(0, 3, 1, 2), b.permute(0, 3, 1, 2)
    return a.contiguous(), b.contiguous()          # [B, C, H, W]


def _ssim_map(a, b, win=7, data_range=2.0, k1=0.01, k2=0.03):
    """skimage.metrics.structural_similarity(full=True), gaussian_weights=False, use_sample_covariance=True."""
    pool = lambda x: torch.nn.functional.avg_pool2d(x, win, stride=1)
    n = win * win
    cov_norm = n / (n - 1.0)
    ux, uy = pool(a), pool(b)
    vx = cov_norm * (pool(a * a) - ux * ux)
    vy = cov_norm * (pool(b * b) - uy * uy)
    vxy = cov_norm * (pool(a * b) - ux * uy)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    return ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))


def compute_img_metric(im1t, im2t, metric="mse", margin=0, mask=None, format=None):
    """utils/metrics.py:18-100 for metric in {"mse", "psnr", "ssim"}; masks as in the reference (mse / psnr: masked images and the
    -10 log10(HW / n_valid) correction; ssim: masked mean of the SSIM map, mask cropped to the map's valid interior)."""
    if metric not in ("mse", "psnr", "ssim"):
        raise RuntimeError(f"img_utils:: metric {metric} not recognized" if metric != "lpips"
                           else "lpips needs the pretrained network: use the reference's utils/metrics.py")
    a, b = _prep(im1t, im2t, format)
    B, _, H, W = a.shape
    m = None
    if mask is not None:
        m = torch.as_tensor(mask, dtype=torch.float32, device=a.device)
        if m.dim() == 3:
            m = m[:, None]
        m = m.expand(-1, a.shape[1], -1, -1) if m.shape[1] == 1 else m
    if margin > 0:
        mh, mw = int(H * margin) + 1, int(W * margin) + 1
        a, b = a[:, :, mh:H - mh, mw:W - mw], b[:, :, mh:H - mh, mw:W - mw]
        m = m[:, :, mh:H - mh, mw:W - mw] if m is not None else None
    vals = []
    for i in range(B):
        x, y = a[i], b[i]
        if metric in ("mse", "psnr"):
            if m is not None:
                x, y = x * m[i], y * m[i]
            mse = img2mse(x.contiguous(), y.contiguous()) if x.is_cuda else torch.mean((x - y) ** 2)
            v = float(mse) if metric == "mse" else 10.0 * math.log10(4.0 / float(mse))
            if m is not None:
                v = v - 10.0 * math.log10(x.shape[-2] * x.shape[-1] / float(m[i, 0].sum()))
        else:
            smap = _ssim_map(x[None], y[None])[0]
            if m is not None:
                mm = m[i][:, 3:-3, 3:-3]
                v = float((smap * mm).sum() / mm.sum())
            else:
                v = float(smap.mean())
        vals.append(v)
    return sum(vals) / len(vals)
