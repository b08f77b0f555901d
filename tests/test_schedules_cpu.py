"""Host logic of the training-loop caller against values produced by the unmodified reference
(oracle/make_golden_schedules.py -> tests/golden/case8_*.json): loss-weight schedules, optimizer parameter groups,
checkpoint payload."""
import json
import os

import pytest
import torch

from util import GOLDEN

from evdeblurnerf_b200 import schedules as sch


def test_schedules_match_reference_values():
    cases = json.load(open(os.path.join(GOLDEN, "case8_schedules.json")))
    assert len(cases) >= 14
    for c in cases:
        if c["kind"] == "anneal":
            f = sch.annealing_interpolator(c["start_value"], c["end_value"], c["end_step"], c["method"], start_step=c["start_step"])
            for s, v in zip(c["steps"], c["values"]):
                assert f(s) == pytest.approx(v, rel=1e-12, abs=1e-15), (c["method"], s)
        else:
            for i, v in zip(c["iters"], c["values"]):
                got = sch.exponential_scale_fine_loss_weight(c["N_iters"], c["kernel_start_iter"], c["start_ratio"], c["end_ratio"], i)
                assert got == pytest.approx(v, rel=1e-12)
    with pytest.raises(ValueError):
        sch.annealing_interpolator(0., 1., 10, "quadratic")


def test_linear_schedule_keeps_the_reference_offset_quirk():
    f = sch.annealing_interpolator(0.0, 1.0, 2000, "linear", start_step=1000)      # utils/misc.py:35-36: slope * step, not (step - start)
    assert f(999) == 0.0 and f(1000) == pytest.approx(1.0) and f(1500) == pytest.approx(1.5) and f(2000) == 1.0


@pytest.mark.parametrize("key,mode,wd", [("c2f_wd0", "c2f", 0.0), ("c2f_wd1", "c2f", 1e-4), ("nerf_wd0", "nerf", 0.0)])
def test_optimizer_groups_match_reference_order(key, mode, wd):
    lay = json.load(open(os.path.join(GOLDEN, "case8_optimizer_layout.json")))[key]
    groups = sch.optimizer_groups(lay["named_parameters"], lay["crf_named_parameters"], mode=mode, colornet_weightdecay=wd)
    want = [[n[4:] if n.startswith("crf.") else n for n in g] for g in lay["groups"]]
    assert groups == want


def test_checkpoint_payload_is_a_loadable_adam_state_dict_and_round_trips():
    lay_all = json.load(open(os.path.join(GOLDEN, "case8_optimizer_layout.json")))
    lay = lay_all["c2f_wd0"]
    g = torch.Generator().manual_seed(0)
    shapes = {}
    flat_names = [n for grp in lay["groups"] for n in grp]
    for pid, n in enumerate(flat_names):
        shapes[n] = lay_all["adam_state_shapes"][str(pid)]
    net = {n: torch.randn(shapes[n], generator=g) for n in lay["named_parameters"]}
    net["awpnet.MAM.Corr.convd.1.running_mean"] = torch.zeros(32)             # buffers travel in network_state_dict too
    crf = {n: torch.randn(shapes["crf." + n], generator=g) for n in lay["crf_named_parameters"]}
    m = {n: torch.randn(shapes[n], generator=g) for n in flat_names}
    v = {n: torch.rand(shapes[n], generator=g) for n in flat_names}
    groups = sch.optimizer_groups(lay["named_parameters"], lay["crf_named_parameters"])
    ck = sch.checkpoint_dict(1234, net, crf, m, v, groups, lr=4e-4, initial_lr=5e-4)
    assert set(ck) == {"wandb_id", "global_step", "crf_state_dict", "network_state_dict", "optimizer_state_dict"}   # run_nerf.py:628-634
    assert sorted(ck["optimizer_state_dict"]["param_groups"][0].keys()) == lay_all["adam_param_group_keys"]
    assert sorted(ck["optimizer_state_dict"]["state"][0].keys()) == lay_all["adam_state_keys"]
    # a torch.optim.Adam built the reference way (run_nerf.py:243-266) accepts it and ends up with our moments at the right parameters
    params = {n: torch.nn.Parameter(torch.zeros(shapes[n])) for n in flat_names}
    opt = torch.optim.Adam([{"params": [params[n] for n in grp], "lr": 5e-4} for grp in lay["groups"]], lr=5e-4, betas=(0.9, 0.999))
    opt.load_state_dict(ck["optimizer_state_dict"])
    for n in ("mlp_fine.color_net.1.weight", "mlp_coarse.app_plane.2", "kernelsnet.r_linear.bias", "crf.tonemapping_event.linear.6.bias"):
        assert torch.equal(opt.state[params[n]]["exp_avg"], m[n]) and torch.equal(opt.state[params[n]]["exp_avg_sq"], v[n]), n
        assert float(opt.state[params[n]]["step"]) == 1234.0
    assert opt.param_groups[0]["lr"] == 4e-4 and opt.param_groups[0]["initial_lr"] == 5e-4
    # ... and what that optimizer saves comes back to names
    step, net2, crf2, m2, v2, adam_step = sch.load_checkpoint_dict({**ck, "optimizer_state_dict": opt.state_dict()}, groups)
    assert step == 1234 and adam_step == 1234 and set(net2) == set(net) and set(crf2) == set(crf)
    assert all(torch.equal(m2[n], m[n]) and torch.equal(v2[n], v[n]) for n in flat_names)
    with pytest.raises(ValueError):
        sch.load_checkpoint_dict(ck, groups[:-2] + [groups[-2] + ["extra"], groups[-1]])
