"""Host-side schedules and the checkpoint layout of the reference training loop.

* `annealing_interpolator`, `exponential_scale_fine_loss_weight`: utils/misc.py:9-56 -- loss-weight schedules used at
  run_nerf.py:121-142 (event EGM weight, pts0-prior weight, kernel warm-up) and :466-471 (AWP coarse-to-fine mix).
* `optimizer_groups`, `checkpoint_dict`, `load_checkpoint_dict`: the parameter groups of run_nerf.py:243-261 and the `.tar`
  payload of run_nerf.py:628-634 / its reload at :282-295, so that checkpoints move between the two code bases.

Pure host logic: nothing here touches the GPU."""
import math
import re

import torch

COLOR_WEIGHT_RE = r"\.color_net\.[0-9]+\.weight"          # run_nerf.py:246


def annealing_interpolator(start_value, end_value, end_step, method="linear", start_step=0):
    """step -> weight.  `linear` keeps the reference's arithmetic exactly: the ramp is `start + slope * step`, NOT
    `slope * (step - start_step)` (utils/misc.py:35-36), so with start_step > 0 it jumps at start_step."""
    if method == "linear":
        def linear(step):
            if step >= end_step:
                return end_value
            if step < start_step:
                return start_value
            return start_value + (end_value - start_value) / (end_step - start_step) * step
        return linear
    if method == "cosine":
        def cosine(step):
            if step >= end_step:
                return end_value
            if step < start_step:
                return start_value
            c = (1 + math.cos(math.pi * (step - start_step) / (end_step - start_step))) / 2
            return start_value * c + end_value * (1 - c)
        return cosine
    if method == "constant":
        return lambda step: start_value
    raise ValueError("Unsupported method: {}".format(method))


def exponential_scale_fine_loss_weight(N_iters, kernel_start_iter, start_ratio, end_ratio, iter):
    """utils/misc.py:9-12: start_ratio * exp(log(end/start) * (iter - kernel_start) / (N_iters - kernel_start))."""
    scale = (1 / (N_iters - kernel_start_iter)) * math.log(end_ratio / start_ratio)
    return start_ratio * math.exp(scale * (iter - kernel_start_iter))


def _is_vol(name):
    return "app_plane" in name or "app_line" in name


def optimizer_groups(names, crf_names=(), mode="c2f", colornet_weightdecay=0.0):
    """Parameter names per torch.optim.Adam group, in the order run_nerf.py:243-261 builds them.  `names`: the trainable network
    parameter names in `nerf.named_parameters()` order (= their order in the reference state_dict).
      c2f, weight decay : [color_net.N.weight (named order)], [other non-VM (named order)], grad_vars_vol, crf
      c2f               : grad_vars, grad_vars_vol, crf     (renderer.py:60-79, voxnerf.py:120-124)
      nerf              : nerf.parameters(), crf"""
    names = list(names)
    if mode == "nerf":
        return [names, list(crf_names)]

    def field(prefix):
        mine = [n for n in names if n.startswith(prefix)]
        vol = [n for n in mine if "app_line" in n] + [n for n in mine if "app_plane" in n]
        net = [n for n in mine if ".basis_mat." in n] + [n for n in mine if ".color_net." in n] + [n for n in mine if ".sigma_net." in n]
        return vol, net
    vol_c, net_c = field("mlp_coarse.")
    vol_f, net_f = field("mlp_fine.")
    vol = vol_c + vol_f
    if colornet_weightdecay:
        net_named = [n for n in names if not _is_vol(n)]
        groups = [[n for n in net_named if re.findall(COLOR_WEIGHT_RE, n)], [n for n in net_named if not re.findall(COLOR_WEIGHT_RE, n)], vol]
    else:
        net = net_c + [n for n in names if n.startswith("kernelsnet.")] + [n for n in names if n.startswith("awpnet.")] + net_f
        groups = [net, vol]
    return groups + [list(crf_names)]


def checkpoint_dict(global_step, network_state, crf_state, exp_avg, exp_avg_sq, groups, lr, initial_lr, colornet_weightdecay=0.0,
                    adam_step=None, wandb_id=None):
    """The payload run_nerf.py:628-634 saves.  `exp_avg` / `exp_avg_sq`: name -> tensor (crf entries under 'crf.<name>');
    `groups` from `optimizer_groups` (crf names WITHOUT the 'crf.' prefix).  The optimizer entry is a genuine
    torch.optim.Adam state_dict (integer parameter ids in group order), loadable by the reference at run_nerf.py:295."""
    state, param_groups, pid = {}, [], 0
    step_t = torch.tensor(float(global_step if adam_step is None else adam_step))
    for gi, group in enumerate(groups):
        ids = []
        is_crf = gi == len(groups) - 1
        for n in group:
            key = ("crf." + n) if is_crf else n
            if key in exp_avg:
                state[pid] = {"step": step_t.clone(), "exp_avg": exp_avg[key].detach().clone().cpu(),
                              "exp_avg_sq": exp_avg_sq[key].detach().clone().cpu()}
            ids.append(pid)
            pid += 1
        param_groups.append({"lr": lr, "weight_decay": colornet_weightdecay if (colornet_weightdecay and gi == 0) else 0,
                             "initial_lr": initial_lr, "betas": (0.9, 0.999), "eps": 1e-8, "amsgrad": False, "maximize": False,
                             "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                             "decoupled_weight_decay": False, "params": ids})
    return {"wandb_id": wandb_id, "global_step": int(global_step),
            "crf_state_dict": {k: v.detach().clone().cpu() for k, v in crf_state.items()},
            "network_state_dict": {k: v.detach().clone().cpu() for k, v in network_state.items()},
            "optimizer_state_dict": {"state": state, "param_groups": param_groups}}


def load_checkpoint_dict(ckpt, groups):
    """Inverse of `checkpoint_dict` (also accepts checkpoints the reference wrote): -> (global_step, network_state, crf_state,
    exp_avg {name: tensor}, exp_avg_sq {name: tensor}, adam_step)."""
    opt = ckpt.get("optimizer_state_dict") or {"state": {}, "param_groups": []}
    pgs = opt["param_groups"]
    if pgs and [len(g["params"]) for g in pgs] != [len(g) for g in groups]:
        raise ValueError(f"optimizer groups {[len(g['params']) for g in pgs]} do not match this model's {[len(g) for g in groups]}")
    m, v, step = {}, {}, None
    for gi, (pg, group) in enumerate(zip(pgs, groups)):
        is_crf = gi == len(groups) - 1
        for pid, n in zip(pg["params"], group):
            st = opt["state"].get(pid)
            if st is None:
                continue
            key = ("crf." + n) if is_crf else n
            m[key], v[key] = st["exp_avg"], st["exp_avg_sq"]
            step = int(st["step"]) if step is None else step
    return int(ckpt["global_step"]), ckpt["network_state_dict"], ckpt.get("crf_state_dict", {}), m, v, step
