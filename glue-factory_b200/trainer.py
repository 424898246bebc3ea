"""Data-parallel training step around the matcher (one process per GPU).

Restates the per-iteration body of the reference trainer
(/root/reference/gluefactory/train.py:456-517: zero_grad -> forward -> loss ->
mean -> backward -> gradient exchange -> Adam) with the two B200-specific
changes SURVEY.md section 2.2 (C1) calls for:

  * every parameter lives in ONE flat fp32 buffer and every gradient in one flat
    gradient buffer (the nn.Parameters are views), so the gradient exchange is a
    single `all_reduce` over NVLink per optimiser step instead of DDP's bucketed
    hooks plus the extra scalar all-reduce of train.py:483-488;
  * Adam runs as one kernel over the flat buffers (`lgb200_adam_flat`), with the
    1/world averaging folded into it.

The matcher is per-pair, so image pairs shard across ranks with no other
collective.
"""
import torch
import torch.distributed as dist

from . import ops
from .synthetic import to_device

_ALIGN = 64  # elements; keeps every parameter view 256-byte aligned


class FlatParams:
    """Re-homes the trainable parameters of `module` into one flat buffer (+ flat grad buffer)."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev, dt = self.params[0].device, torch.float32
        self.offsets, n = [], 0
        for p in self.params:
            assert p.dtype == dt and p.device == dev
            self.offsets.append(n)
            n += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = n
        self.flat = torch.zeros(n, device=dev, dtype=dt)
        self.grad = torch.zeros(n, device=dev, dtype=dt)
        for p, off in zip(self.params, self.offsets):
            self.flat[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + p.numel()].view_as(p)
            p.grad = self.grad[off:off + p.numel()].view_as(p)
        # bf16 shadow of the whole buffer (one cast kernel per step) for the tensor-core GEMMs
        self._names = {id(p): n for n, p in module.named_parameters()}
        self.flat16 = None
        self._views16 = None
        self.grad_views = [self.grad[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]
        module._b200_flat = self

    def shadow_bf16(self):
        """name -> bf16 view of the parameter, refreshed from the fp32 master copy by ONE cast kernel."""
        if self.flat16 is None:
            self.flat16 = torch.empty(self.numel, device=self.flat.device, dtype=torch.bfloat16)
            self._views16 = {self._names[id(p)]: self.flat16[off:off + p.numel()].view_as(p)
                             for p, off in zip(self.params, self.offsets)}
        ops.call("lgb200_cast_bf16", ops.ptr(self.flat), ops.ptr(self.flat16), self.numel, ops.stream_ptr())
        return self._views16

    def zero_grad(self):
        """Detach the .grad views: autograd then hands each freshly computed gradient over by reference
        (no per-parameter accumulate kernel); `gather_grads` packs them into the flat buffer."""
        for p in self.params:
            p.grad = None

    def gather_grads(self):
        have = [(v, p.grad) for v, p in zip(self.grad_views, self.params) if p.grad is not None]
        missing = [v for v, p in zip(self.grad_views, self.params) if p.grad is None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if missing:
            torch._foreach_zero_(missing)
        for p, v in zip(self.params, self.grad_views):
            p.grad = v


class MatcherTrainer:
    """model: a matcher with forward(data)->pred and loss(pred, data)->(losses, metrics)."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None):
        self.model = model.train()
        self.fp = FlatParams(model)
        self.m = torch.zeros_like(self.fp.flat)
        self.v = torch.zeros_like(self.fp.flat)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.t = 0
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1

    def exchange_gradients(self):
        """The ONE collective of the step: sum-all-reduce of the flat gradient buffer."""
        if self.world > 1:
            dist.all_reduce(self.fp.grad, op=dist.ReduceOp.SUM, group=self.group)

    # ---- CUDA-graph replay of the whole step (removes the ~2500 per-step launch calls from the host)
    def capture(self, example, device, warmup=3):
        """Capture forward + loss + backward + all-reduce + Adam into one CUDA graph.  `example` fixes the
        shapes; later batches are copied into the captured static input buffers."""
        # private static inputs: to_device() returns device tensors as they are, and step_graphed() overwrites these
        self._static = _clone(to_device(example, device))
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_impl(self._static)
        torch.cuda.current_stream().wait_stream(side)
        self._t_dev = torch.full((1,), self.t, device=device, dtype=torch.int32)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss, self._static_losses = self._step_impl(self._static, graphed=True)
        return self

    def step_graphed(self, data, prefetch=None):
        """Replay the captured step on a new batch (host or device tensors of the captured shapes).

        `prefetch`: the batch of the NEXT call (pinned host memory).  Its host-to-device copy is queued on a side
        stream into a staging set right after this step's graph launch, so it overlaps this step's compute; the next
        call (which must pass that same object as `data`) then only does a device-to-device copy into the captured
        static inputs."""
        cur = torch.cuda.current_stream()
        if getattr(self, "_staged_src", None) is data and data is not None:
            cur.wait_event(self._staged_ready)
            _copy_into(self._static, self._staged)
            self._staged_src = None
            self._staged_free.record(cur)
        else:
            _copy_into(self._static, data)
        self._graph.replay()
        if prefetch is not None:
            if getattr(self, "_staged", None) is None:
                self._staged = _empty_like(self._static)
                self._copy_stream = torch.cuda.Stream(device=self.fp.flat.device)
                self._staged_ready, self._staged_free = torch.cuda.Event(), torch.cuda.Event()
                self._staged_free.record(cur)
            self._copy_stream.wait_event(self._staged_free)  # the last D2D out of the staging set has finished
            with torch.cuda.stream(self._copy_stream):
                _copy_into(self._staged, prefetch)
                self._staged_ready.record(self._copy_stream)
            self._staged_src = prefetch
        self.t += 1
        return self._static_loss, self._static_losses

    def step(self, data, device=None):
        """One optimiser step on one batch; `data` may live in (pinned) host memory."""
        if device is not None:
            data = to_device(data, device, non_blocking=True)
        return self._step_impl(data)

    def _step_impl(self, data, graphed=False):
        self.fp.zero_grad()
        pred = self.model(data)
        losses, _ = self.model.loss(pred, data)
        loss = losses["total"].mean()
        loss.backward()
        self.fp.gather_grads()
        self.exchange_gradients()
        if graphed:
            # the bias-correction terms depend on the step count, which must not be baked into the graph:
            # the count lives in device memory and is incremented inside the graph.
            self._t_dev.add_(1)
            ops.adam_flat_(self.fp.flat, self.fp.grad, self.m, self.v, 0, self.lr, self.betas, self.eps, self.wd,
                           grad_scale=1.0 / self.world, step_dev=self._t_dev)
        else:
            self.t += 1
            ops.adam_flat_(self.fp.flat, self.fp.grad, self.m, self.v, self.t, self.lr, self.betas, self.eps, self.wd,
                           grad_scale=1.0 / self.world)
        return loss.detach(), losses


def _clone(tree):
    return {k: (_clone(v) if isinstance(v, dict) else v.clone() if torch.is_tensor(v) else v) for k, v in tree.items()}


def _empty_like(tree):
    return {k: (_empty_like(v) if isinstance(v, dict) else torch.empty_like(v) if torch.is_tensor(v) else v)
            for k, v in tree.items()}


def _copy_into(dst, src):
    for k, v in dst.items():
        if isinstance(v, dict):
            _copy_into(v, src[k])
        elif torch.is_tensor(v):
            v.copy_(src[k], non_blocking=True)
