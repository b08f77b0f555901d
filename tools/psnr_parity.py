#!/usr/bin/env python
"""Image-quality parity of the bf16 tensor-core mode against the fp32 parity mode on a synthetic scene (no dataset or
checkpoint is available offline): a hidden "teacher" model (random VM grids + MLPs + blur kernel) renders blurred target
colours for a fixed ray set; two students with identical initialisation and batches are trained with `Trainer`, one per
precision, and evaluated on held-out rays (each student both in its own mode and in the fp32 mode).

    python tools/psnr_parity.py [--steps 400] [--rays 4096]  -> JSON lines (profiles/r1_psnr_parity_synthetic.jsonl)
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

E = 5


def make_model(seed, coarse=(48, 48, 32), fine=(96, 96, 64), plane_scale=0.1, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    P = {}

    def lin(name, o, i):
        b = 1.0 / (i ** 0.5)
        P[name] = ((torch.rand(o, i, generator=g) * 2 - 1) * b).to(device)

    for pre, gs, hid, geo in (("mlp_coarse.", coarse, 64, 15), ("mlp_fine.", fine, 256, 128)):
        for i, (m, v) in enumerate((((0, 1), 2), ((0, 2), 1), ((1, 2), 0))):
            c = (64, 16, 16)[i]
            P[pre + f"app_plane.{i}"] = (plane_scale * torch.randn(1, c, gs[m[1]], gs[m[0]], generator=g)).to(device)
            P[pre + f"app_line.{i}"] = (plane_scale * torch.randn(1, c, gs[v], 1, generator=g)).to(device)
        lin(pre + "basis_mat.weight", 32, 96)
        lin(pre + "sigma_net.0.weight", hid, (32 if hid == 64 else 64) + 63)
        lin(pre + "sigma_net.1.weight", 1 + geo, hid)
        lin(pre + "color_net.0.weight", hid, geo + 27)
        lin(pre + "color_net.1.weight", hid, hid)
        lin(pre + "color_net.2.weight", 3, hid)
    pre = "kernelsnet."
    P[pre + "view_embed_module.img_embed"] = (0.5 * torch.randn(bench.N_IMGS, 32, generator=g)).to(device)
    for h in ("r", "v", "w"):
        lin(pre + f"{h}_branch.0.weight", 32, 32)
        P[pre + f"{h}_branch.0.bias"] = torch.zeros(32, device=device)
    for h, n_out in (("r", 3 * (E - 1)), ("v", 3 * (E - 1)), ("w", E)):
        P[pre + f"{h}_linear.weight"] = (0.05 * torch.randn(n_out, 32, generator=g)).to(device)
        P[pre + f"{h}_linear.bias"] = torch.zeros(n_out, device=device)
    return P


def psnr(a, b):
    return -10.0 * math.log10(float(torch.mean((a - b) ** 2)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--pool", type=int, default=32768)
    ap.add_argument("--eval-every", type=int, default=100)
    a = ap.parse_args()
    from evdeblurnerf_b200 import NeRFAll
    from evdeblurnerf_b200.trainer import Trainer
    dev = "cuda"
    rk = dict(N_samples=64, N_importance=64, perturb=0., raw_noise_std=0.)
    tp = make_model(1, plane_scale=1.0)
    for pre in ("mlp_coarse.", "mlp_fine."):        # a contrasty teacher: strong colour and density heads
        tp[pre + "color_net.2.weight"] *= 12.0
        tp[pre + "sigma_net.1.weight"][0] *= 40.0
    teacher = NeRFAll(tp, *bench.AABB, kernel_ptnum=E, precision="fp32").eval()
    rays, idx = bench.make_rays(a.pool + 8192, seed=5)
    rays, idx = rays.to(dev), idx.to(dev)
    with torch.no_grad():
        target = torch.cat([teacher.render_blurred(bench.H, bench.W, bench.KMAT, rays[i:i + 8192], idx[i:i + 8192], **rk)[0]
                            for i in range(0, rays.shape[0], 8192)])
    tr_rays, tr_idx, tr_t = rays[:a.pool], idx[:a.pool], target[:a.pool]
    ev_rays, ev_idx, ev_t = rays[a.pool:], idx[a.pool:], target[a.pool:]
    print(json.dumps({"event": "setup", "train_rays": a.pool, "eval_rays": int(ev_rays.shape[0]), "target_mean": float(target.mean()),
                      "target_std": float(target.std())}), flush=True)
    curves = {}
    for precision in ("fp32", "bf16"):
        init = make_model(2)                                                    # same initialisation for both students
        tr = Trainer(init, None, *bench.AABB, kernel_ptnum=E, precision=precision, lrate=2e-3, lrate_decay=250, tv_loss_weight=0.0,
                     crf_kwargs=dict(map_type_rgb="none", map_type_event="gamma"),
                     render_kwargs=dict(N_samples=64, N_importance=64, perturb=1., raw_noise_std=0.), seed=3)
        perm = torch.Generator().manual_seed(11)
        curve = []
        for step in range(a.steps + 1):
            if step % a.eval_every == 0:
                state = tr.state_dict()
                row = {"event": "eval", "precision": precision, "step": step}
                for mode in sorted({precision, "fp32"}):
                    student = NeRFAll(state, *bench.AABB, kernel_ptnum=E, precision=mode).eval()
                    with torch.no_grad():
                        out = student.render_blurred(bench.H, bench.W, bench.KMAT, ev_rays, ev_idx, **rk)[0]
                    row[f"psnr_eval_in_{mode}"] = psnr(out, ev_t)
                curve.append(row)
                print(json.dumps(row), flush=True)
            if step == a.steps:
                break
            sel = torch.randint(0, a.pool, (a.rays,), generator=perm).to(dev)
            tr.step({"rays": tr_rays[sel], "images_idx": tr_idx[sel], "rgbsf": tr_t[sel]}, bench.H, bench.W, bench.KMAT)
        curves[precision] = curve
        del tr
        torch.cuda.empty_cache()
    f32, b16 = curves["fp32"][-1], curves["bf16"][-1]
    print(json.dumps({"event": "summary", "steps": a.steps, "psnr_fp32_student": f32["psnr_eval_in_fp32"],
                      "psnr_bf16_student_eval_bf16": b16["psnr_eval_in_bf16"], "psnr_bf16_student_eval_fp32": b16["psnr_eval_in_fp32"],
                      "delta_db": b16["psnr_eval_in_bf16"] - f32["psnr_eval_in_fp32"]}), flush=True)


if __name__ == "__main__":
    main()
