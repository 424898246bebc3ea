"""GPU: the flat fused optimiser step -- Adam + GradScaler unscale / inf check / skip + LR groups on
the device, against torch.optim.Adam + torch.amp.GradScaler (train.py:347-367, 456, 490-517), and the trainer's
NaN guard (train.py:477-480, 503-512), LR scheduling under CUDA-graph replay and side-effect-free capture."""
import numpy as np
import pytest
import torch

from gluefactory_b200 import ops, synthetic
from gluefactory_b200.matchers.lightglue import LightGlue
from gluefactory_b200.trainer import MatcherTrainer
from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, generator=g).to(DEV)


def test_flat_adam_with_grad_scaler_matches_torch():
    """Same gradients through (a) torch.optim.Adam + torch.amp.GradScaler and (b) flat_grad_check -> amp_update ->
    adam_flat with device-resident step / scale / found_inf: identical parameters, identical scale trajectory,
    the overflow step skipped by both."""
    n = 50_003
    p = _rand(n, 1)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=2e-3, weight_decay=0.01)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, device=DEV, dtype=torch.int32)
    scale = torch.full((1,), 1024.0, device=DEV)
    growth = torch.zeros(1, device=DEV, dtype=torch.int32)
    found = torch.zeros(1, device=DEV)
    lr_dev = torch.full((1,), 2e-3, device=DEV)
    scaler.scale(torch.ones(1, device=DEV))  # GradScaler creates its scale tensor lazily on the first scale() call
    for t in range(6):
        g = _rand(n, 10 + t)
        if t == 2:
            g[777] = float("inf")  # overflow: both must skip and halve the scale
        cur = scaler.get_scale()
        ref.grad = g.clone() * cur           # what backward of (loss * scale) leaves in .grad
        scaler.step(opt)
        scaler.update()
        gs = g * float(scale.item())
        ops.flat_grad_check(gs, found)
        ops.amp_update(found, step_dev=step)                       # ++step unless skipped
        ops.adam_flat_(p, gs, m, v, 0, 0.0, weight_decay=0.01, step_dev=step, lr_dev=lr_dev, loss_scale_dev=scale,
                       found_inf_dev=found)                        # un-scales with the scale the gradients carry
        ops.amp_update(found, None, scale, growth, 2.0, 0.5, 2)    # GradScaler.update
        assert float(scale.item()) == scaler.get_scale(), (t, float(scale.item()), scaler.get_scale())
    assert int(step.item()) == 5
    assert rel_err(p, ref.data) < 1e-6


def _model(L=2, seed=51):
    conf = dict(synthetic.DEFAULT_CONF, n_layers=L)
    model = LightGlue(dict(conf, precision="bf16"))
    model.load_state_dict({k: v.float() for k, v in synthetic.make_weights(conf, seed=seed).items()}, strict=False)
    return model.to(DEV).train()


@pytest.mark.parametrize("graphed", [False, True])
def test_non_finite_batch_is_skipped_on_device(graphed):
    """train.py:477-480: a NaN loss must not touch the parameters or the Adam state; the next clean batch trains."""
    good = synthetic.to_device(synthetic.make_pairs(2, 256, seed=60), DEV)
    bad = synthetic.to_device(synthetic.make_pairs(2, 256, seed=61), DEV)
    bad["descriptors0"][1, 3, 5] = float("nan")
    tr = MatcherTrainer(_model(), lr=1e-3)
    if graphed:
        tr.capture(good, DEV)
    run = tr.step_graphed if graphed else tr.step
    run(good)
    assert tr.t == 1 and not tr.skipped_last_step()
    before = (tr.fp.flat.clone(), tr.m.clone(), tr.v.clone())
    loss, _ = run(bad)
    assert not torch.isfinite(loss)
    assert tr.skipped_last_step() and tr.t == 1
    assert torch.equal(tr.fp.flat, before[0]) and torch.equal(tr.m, before[1]) and torch.equal(tr.v, before[2])
    loss, _ = run(good)
    assert torch.isfinite(loss) and tr.t == 2 and not tr.skipped_last_step()
    assert not torch.equal(tr.fp.flat, before[0]) and torch.isfinite(tr.fp.flat).all()


def test_capture_leaves_no_trace_and_lr_is_live_under_replay():
    data = synthetic.to_device(synthetic.make_pairs(2, 256, seed=62), DEV)
    tr = MatcherTrainer(_model(), lr=1e-3)
    p0 = tr.fp.flat.clone()
    tr.capture(data, DEV, warmup=2)
    assert torch.equal(tr.fp.flat, p0) and tr.t == 0 and float(tr.m.abs().sum()) == 0.0
    tr.step_graphed(data)
    d1 = (tr.fp.flat - p0).abs().max().item()
    assert d1 > 0
    # lr = 0 through the scheduler hook: the replayed graph must see it (round-1 graphs baked lr in by value)
    tr.lr = 0.0
    p1 = tr.fp.flat.clone()
    tr.step_graphed(data)
    assert torch.equal(tr.fp.flat, p1)
    tr.lr = 1e-3
    tr.step_graphed(data)
    assert not torch.equal(tr.fp.flat, p1)
    # eager twin: same three effective steps give the same parameters
    te = MatcherTrainer(_model(), lr=1e-3)
    te.step(data)
    te.lr = 0.0
    te.step(data)
    te.lr = 1e-3
    te.step(data)
    assert rel_err(tr.fp.flat, te.fp.flat) < 1e-5


def test_loss_scale_and_lr_groups_match_plain_training():
    """Backward is linear in the incoming gradient (SURVEY 8b 'Autocast / dtype'): training with a 2^12 loss scale
    gives the same update as without; an LR group with factor 0 freezes exactly the matching parameters."""
    data = synthetic.to_device(synthetic.make_pairs(2, 192, seed=63), DEV)
    a = MatcherTrainer(_model(), lr=1e-3)
    b = MatcherTrainer(_model(), lr=1e-3, loss_scale=4096.0)
    for _ in range(2):
        a.step(data)
        b.step(data)
    assert b.loss_scale() == 4096.0 and b.t == 2
    assert rel_err(b.fp.flat, a.fp.flat) < 2e-3  # bf16 gradients of a scaled loss round differently
    c = MatcherTrainer(_model(), lr=1e-3, lr_scaling=[(0.0, ["log_assignment"]), (10.0, ["posenc"])])
    before = {n: p.detach().clone() for n, p in c.model.named_parameters()}
    c.step(data)
    for n, p in c.model.named_parameters():
        moved = not torch.equal(p.detach(), before[n])
        assert moved != ("log_assignment" in n), n
    # the factor-10 group moved 10x as far as it does in the plain trainer's first step (Adam's first step = lr * sign)
    d = MatcherTrainer(_model(), lr=1e-3)
    d.step(data)
    dp_c = (dict(c.model.named_parameters())["posenc.Wr.weight"].detach() - before["posenc.Wr.weight"]).abs().mean()
    dp_d = (dict(d.model.named_parameters())["posenc.Wr.weight"].detach() - before["posenc.Wr.weight"]).abs().mean()
    np.testing.assert_allclose((dp_c / dp_d).item(), 10.0, rtol=1e-3)
