"""CPU tests of the training-step component (raft_optical_flow_b200/train.py; scope table 8f, f4): the sequence
loss against a fixture produced by the reference's own function (tests/golden/make_golden.py::case_sequence_loss)
and the optimiser / schedule / step plumbing on a toy model."""
import numpy as np
import torch

from conftest import load_golden
from raft_optical_flow_b200 import train


def test_sequence_loss_matches_reference_fixture():
    g = load_golden("sequence_loss")
    preds = [torch.from_numpy(p).clone().requires_grad_(True) for p in g["preds"]]
    loss, metrics = train.sequence_loss(preds, torch.from_numpy(g["flow_gt"]), torch.from_numpy(g["valid"]), gamma=0.8)
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    got = train.metrics_to_host(metrics)
    assert np.allclose([got[k] for k in ("epe", "1px", "3px", "5px")], g["metrics"], rtol=1e-5, atol=1e-6)
    loss.backward()
    for p, want in zip(preds, g["dpreds"]):
        assert np.abs(p.grad.numpy() - want).max() <= 1e-6 * max(np.abs(want).max(), 1e-12) + 1e-12


def test_fetch_optimizer_is_adamw_with_linear_one_cycle():
    net = torch.nn.Linear(4, 3)
    opt, sched = train.fetch_optimizer(net, lr=4e-4, wdecay=1e-4, epsilon=1e-8, num_steps=1000)
    assert isinstance(opt, torch.optim.AdamW) and opt.defaults["weight_decay"] == 1e-4 and opt.defaults["eps"] == 1e-8
    lrs = []
    for _ in range(1100):
        opt.step()
        sched.step()
        lrs.append(opt.param_groups[0]["lr"])
    peak = int(np.argmax(lrs))
    assert abs(max(lrs) - 4e-4) < 1e-9 and 50 <= peak <= 60  # 5 % warm-up of num_steps + 100 (train.py:114-116)
    assert lrs[-1] < 1e-6  # annealed linearly to ~0


class _ToyFlow(torch.nn.Module):
    """Stands in for RAFT: returns `iters` predictions [N, 2, H, W]."""

    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(6, 2, 3, padding=1)

    def forward(self, image1, image2, iters=3):
        f = self.conv(torch.cat([image1, image2], 1) / 255.0)
        return [f * (i + 1) / iters for i in range(iters)]


def test_train_step_single_process_equals_reference_loop_body():
    torch.manual_seed(0)
    net, twin = _ToyFlow(), _ToyFlow()
    twin.load_state_dict(net.state_dict())
    im1, im2 = 255 * torch.rand(2, 3, 8, 10), 255 * torch.rand(2, 3, 8, 10)
    gt, valid = torch.randn(2, 2, 8, 10), torch.ones(2, 8, 10)
    step = train.TrainStep(net, num_steps=100, iters=3, clip=1.0)
    # the reference's loop body (train.py:199-228) spelled out on the twin
    opt, sched = train.fetch_optimizer(twin, num_steps=100)
    for _ in range(3):
        loss, metrics = step(im1, im2, gt, valid)
        opt.zero_grad()
        want, _ = train.sequence_loss(twin(im1, im2, iters=3), gt, valid, 0.8)
        want.backward()
        torch.nn.utils.clip_grad_norm_(twin.parameters(), 1.0)
        opt.step()
        sched.step()
        assert abs(loss.item() - want.item()) < 1e-6
    for a, b in zip(net.parameters(), twin.parameters()):
        assert torch.allclose(a, b, atol=1e-7)
    assert step.steps_done == 3 and step.allreduce_launches == 0 and set(metrics) == {"epe", "1px", "3px", "5px"}
    step.close()


def test_bench_arms_describe_the_same_workload():
    """`bench.py --impl ours` and `--impl reference` must print the same `config` dict (the driver compares them)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("rcb_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for name, cfg in bench.CONFIGS.items():
        a = bench.config_dict(name, cfg, "all-pairs")
        b = bench.config_dict(name, cfg, "all-pairs")
        assert a == b and a["workload"].startswith(name) and a["pairs_per_gpu"] == cfg[0]
    build, lookup, flops = bench.algorithmic_bytes(8, 256, 55, 128, 4, 4)
    assert build == 2205941760 and lookup == 163553280 and abs(flops - 2.0 * 8 * 7040 * 7040 * 256) < 1
    assert bench.cpu_sample_pairs(8, 55, 128) == 8 and bench.cpu_sample_pairs(4, 136, 240) == 1
