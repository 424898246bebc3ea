"""GPU: the plugin where the reference actually calls it:

  * inside `torch.autocast(cuda, bfloat16|float16)` with a GradScaler-scaled loss (train.py:468-472, 490);
  * through a TwoViewPipeline-shaped caller (two_view_pipeline.py:80-113: matcher({**data, **pred}), ground truth
    component run inside loss(), `gt_` prefixed labels, matcher.loss(pred, {**pred, **data}));
  * wrapped in DistributedDataParallel at world size 2, `loss_fn` bound before the wrap (train.py:334-339), gradients
    compared with the flat-buffer trainer's single all-reduce.

/root/reference does not exist on the GPU box, so the pipeline is restated here in ~25 lines (the CPU suite constructs
the plugin through the unmodified reference get_model / TwoViewPipeline, tests/test_abi_and_plugin_cpu.py).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gluefactory_b200 import synthetic
from gluefactory_b200.matchers.homography_matcher import HomographyMatcher
from gluefactory_b200.matchers.lightglue import LightGlue
from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(L=2, seed=81, **kw):
    conf = dict(synthetic.DEFAULT_CONF, n_layers=L)
    model = LightGlue(dict(conf, precision="bf16", **kw))
    model.load_state_dict({k: v.float() for k, v in synthetic.make_weights(conf, seed=seed).items()}, strict=False)
    return model.to(DEV).train()


class PipelineLike(torch.nn.Module):
    """two_view_pipeline.py:80-113 with extractor = None (features come with the batch, the cached-feature route
    :62-70), matcher = the plugin, ground_truth = the plugin's homography_matcher, run_gt_in_forward = False."""

    def __init__(self, matcher, ground_truth):
        super().__init__()
        self.matcher, self.ground_truth = matcher, ground_truth

    def forward(self, data):
        pred = {}
        for i in "01":
            pred.update({k + i: v for k, v in data[f"view{i}"]["cache"].items()})
        return {**pred, **self.matcher({**data, **pred})}

    def loss(self, pred, data):
        gt_pred = self.ground_truth({**data, **pred})
        pred.update({f"gt_{k}": v for k, v in gt_pred.items()})
        losses, metrics = self.matcher.loss(pred, {**pred, **data})
        return {**losses, "total": losses["total"] + 0}, metrics


def _pipeline_batch(B, N, seed):
    d = synthetic.to_device(synthetic.make_pairs(B, N, seed=seed, with_gt=False), DEV)
    batch = {
        "H_0to1": d["H_0to1"],
        "view0": {"image_size": d["view0"]["image_size"], "cache": {"keypoints": d["keypoints0"], "descriptors": d["descriptors0"]}},
        "view1": {"image_size": d["view1"]["image_size"], "cache": {"keypoints": d["keypoints1"], "descriptors": d["descriptors1"]}},
    }
    # the same pair with labels from the torch restatement of gt_matches_from_homography on the same fp32 inputs
    asg, m0, m1 = synthetic.gt_matches_from_homography(d["keypoints0"], d["keypoints1"], d["H_0to1"], 3.0, 3.0)
    plain = dict(d, gt_assignment=asg, gt_matches0=m0, gt_matches1=m1)
    return batch, plain


@pytest.mark.parametrize("mp_dtype", [torch.bfloat16, torch.float16])
def test_pipeline_under_autocast_with_grad_scaler(mp_dtype):
    """`python -m gluefactory.train ... --mp bfloat16`: forward + loss inside autocast, scaled backward outside.  The
    plugin's result must not depend on the ambient autocast state, and the scaled backward must be exactly linear."""
    batch, plain = _pipeline_batch(2, 256, seed=82)
    pipe = PipelineLike(_model(), HomographyMatcher({"th_positive": 3.0, "th_negative": 3.0}))
    # labels produced inside loss() by the ground-truth component equal the restated labels of the plain batch
    with torch.autocast("cuda", dtype=mp_dtype):
        pred = pipe(batch)
        losses, _ = pipe.loss(pred, batch)
        loss = losses["total"].mean()
    assert torch.equal(pred["gt_matches0"], plain["gt_matches0"]) and torch.equal(pred["gt_assignment"], plain["gt_assignment"])
    assert pred["log_assignment"].dtype == torch.float32 and pred["matches0"].dtype == torch.int64
    scale = 1024.0
    (loss * scale).backward()
    g_scaled = {n: p.grad.clone() for n, p in pipe.named_parameters()}
    # reference point: same model, no autocast, unscaled loss, labels passed in the batch
    ref = _model()
    pr = ref(plain)
    lr_, _ = ref.loss(pr, plain)
    lr_["total"].mean().backward()
    np.testing.assert_allclose(loss.item(), lr_["total"].mean().item(), rtol=1e-6)
    assert torch.equal(pred["matches0"], pr["matches0"])
    for n, p in ref.named_parameters():
        assert rel_err(g_scaled["matcher." + n] / scale, p.grad) < 1e-5, n


def _ddp_worker(rank, world, port, backend, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank if torch.cuda.device_count() >= world else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from gluefactory_b200.trainer import MatcherTrainer

        data = synthetic.to_device(synthetic.make_pairs(2, 192, seed=90 + rank), dev)
        # (a) reference-style: DDP around the module, loss_fn bound before the wrap, backward through DDP's hooks
        model = _model(seed=83).to(dev)
        loss_fn = model.loss
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index] if backend == "nccl" else None)
        pred = ddp(data)
        losses, _ = loss_fn(pred, data)
        losses["total"].mean().backward()
        g_ddp = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        # (b) the flat-buffer trainer: one all-reduce of the summed gradients, 1/world folded into Adam
        model2 = _model(seed=83).to(dev)
        tr = MatcherTrainer(model2, lr=0.0)
        tr.step(data)
        g_flat = torch.cat([p.grad.reshape(-1) for p in model2.parameters()]) / world
        err = ((g_ddp - g_flat).norm() / g_flat.norm()).item()
        same_across_ranks = g_ddp.clone()
        dist.all_reduce(same_across_ranks, op=dist.ReduceOp.MAX)
        out.put((rank, err, bool(torch.equal(same_across_ranks, g_ddp)), float(g_flat.norm())))
    finally:
        dist.destroy_process_group()


def test_plugin_under_ddp_matches_flat_trainer():
    """world size 2: NCCL when the box has two GPUs, else both ranks share cuda:0 over gloo (the wiring is the same)."""
    world = 2
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, backend, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, err, same, norm in res:
        assert norm > 0 and err < 1e-5, (rank, err)
        assert same, rank
