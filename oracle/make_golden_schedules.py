"""Golden values of the reference's loss-weight schedules, produced by the UNMODIFIED reference functions
(utils/misc.py:9-56) imported from /root/reference in the build container.  Output: tests/golden/case8_schedules.json.
Run:  python oracle/make_golden_schedules.py"""
import json
import os
import sys

sys.path.insert(0, "/root/reference")
from utils.misc import annealing_interpolator, exponential_scale_fine_loss_weight  # noqa: E402

steps = [0, 1, 10, 999, 1000, 1001, 2500, 4999, 5000, 5001, 20000]
cases = []
for method in ("linear", "cosine", "constant"):
    for (a, b, end, start) in ((1.0, 0.1, 5000, 0), (0.0, 1.0, 5000, 1000), (0.1, 1.0, 20000, 0), (0.0, 1.0, 1201, 1200)):
        f = annealing_interpolator(a, b, end, method, start_step=start)
        cases.append({"kind": "anneal", "method": method, "start_value": a, "end_value": b, "end_step": end, "start_step": start,
                      "steps": steps, "values": [float(f(s)) for s in steps]})
for (N, k0) in ((200000, 1200), (30000, 0)):
    its = [k0, k0 + 10000, N // 2, N - 1]
    cases.append({"kind": "exp", "N_iters": N, "kernel_start_iter": k0, "start_ratio": 0.1, "end_ratio": 0.9, "iters": its,
                  "values": [float(exponential_scale_fine_loss_weight(N, k0, 0.1, 0.9, i)) for i in its]})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "case8_schedules.json")
json.dump(cases, open(out, "w"), indent=1)
print("wrote", out, len(cases), "cases")

# ---- parameter groups + checkpoint payload of the reference (run_nerf.py:243-261, 628-634) -------------------------------
# Built from the reference's own NeRFAll / CRF on small grids; records which NAMED parameter sits at which optimizer index,
# and one real torch.optim.Adam state_dict after a step (shapes only matter for the layout test).
import torch  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import reference_harness as rh  # noqa: E402

layout = {}
for mode in ("c2f", "nerf"):
    for wd in (0.0, 1e-4):
        if mode == "nerf" and wd:
            continue
        args = rh.blurfactory_args(mode=mode, coarse_n_voxels=4096, fine_n_voxels=8192)
        nerf, crf = rh.build_reference(args)
        ids = {id(p): n for n, p in nerf.named_parameters()}
        ids.update({id(p): "crf." + n for n, p in crf.named_parameters()})
        if mode == "c2f":                                     # run_nerf.py:243-256, restated on the reference objects
            if wd:
                groups = [nerf.get_parameters("net", match_re=r"\.color_net\.[0-9]+\.weight"),
                          nerf.get_parameters("net", not_match_re=r"\.color_net\.[0-9]+\.weight"), nerf.grad_vars_vol]
            else:
                groups = [nerf.grad_vars, nerf.grad_vars_vol]
        else:
            groups = [list(nerf.parameters())]
        groups = groups + [list(crf.parameters())]
        layout[f"{mode}_wd{int(bool(wd))}"] = {
            "named_parameters": [n for n, _ in nerf.named_parameters()],
            "crf_named_parameters": [n for n, _ in crf.named_parameters()],
            "groups": [[ids[id(p)] for p in g] for g in groups]}
        if mode == "c2f" and not wd:
            opt = torch.optim.Adam([{"params": g, "lr": 5e-4} for g in groups], lr=5e-4, betas=(0.9, 0.999))
            for g in opt.param_groups:
                g.setdefault("initial_lr", g["lr"])
            for p in nerf.parameters():
                p.grad = torch.full_like(p, 0.5)
            for p in crf.parameters():
                p.grad = torch.full_like(p, 0.25)
            opt.step()
            sd = opt.state_dict()
            layout["adam_param_group_keys"] = sorted(sd["param_groups"][0].keys())
            layout["adam_state_keys"] = sorted(sd["state"][0].keys())
            layout["adam_state_shapes"] = {str(k): list(v["exp_avg"].shape) for k, v in sd["state"].items()}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "case8_optimizer_layout.json")
json.dump(layout, open(out, "w"), indent=1)
print("wrote", out)
