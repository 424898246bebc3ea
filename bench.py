#!/usr/bin/env python
"""Headline benchmark: image-pairs/sec, LightGlue matcher TRAINING step (forward + loss + backward +
gradient all-reduce + Adam) on synthetic N=2048, d=256, 9-layer pairs (BASELINE.json configs[2],
"SuperPoint + LightGlue MegaDepth-shape synthetic pairs"), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores (oracle port)

Prints ONE JSON line on rank 0 (contract: see the task statement / DESIGN.md section "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "image-pairs/sec training SuperPoint+LightGlue N=2048 at 1/2/4/8 B200"
UNIT = "image-pairs/s"
N_KPTS, D_DESC, N_LAYERS, N_HEADS = 2048, 256, 9, 4
# algorithmic attention FLOPs per pair, forward + backward, no recompute counted:
# C = 2 N^2 D ; fwd 7C/layer, bwd 14C/layer
ATTN_FLOPS_PER_PAIR = 21 * (2 * N_KPTS * N_KPTS * D_DESC) * N_LAYERS


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (one `nvidia-smi -lms` process
    started before and stopped after it)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            time.sleep(0.15)  # let the first samples arrive before the timed region starts
        except Exception:
            self.proc = None
        self.t0 = time.time()
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        self.rows = [[c.strip() for c in line.split(",")] for line in out.strip().splitlines() if line.strip()]

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4), ("sw_power_cap", 5)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def host_threads():
    """Threads the CPU leg may really use: affinity mask and cgroup CPU quota, capped at 32 (the oracle's
    N=2048 GEMMs are small; more threads only add synchronisation on a many-core host)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return max(1, min(n, int(os.environ.get("LGB200_CPU_THREADS", "32"))))


def pin(data):
    if isinstance(data, dict):
        return {k: pin(v) for k, v in data.items()}
    return data.pin_memory() if torch.is_tensor(data) else data


def nbytes(data):
    if isinstance(data, dict):
        return sum(nbytes(v) for v in data.values())
    return data.numel() * data.element_size() if torch.is_tensor(data) else 0


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm for this path on the host CPU cores.  The reference is
    Python and does not travel to the GPU box, so this times the oracle port (oracle/lightglue_oracle.py,
    pinned to the reference by tests/golden) -- forward + loss + backward + Adam, fp32, all host threads.
    A 'step' is a bounded sample of the workload: ONE pair of the same N=2048 / 9-layer configuration."""
    if rank != 0:
        return
    from gluefactory_b200 import synthetic
    from oracle import lightglue_oracle as O

    torch.set_num_threads(host_threads())
    conf = dict(synthetic.DEFAULT_CONF)
    w = {k: v.clone().requires_grad_(True) for k, v in synthetic.make_weights(conf, seed=0).items()}
    data = synthetic.make_pairs(1, N_KPTS, seed=1234)
    budget = float(os.environ.get("LGB200_REF_BUDGET_S", "240"))
    t0 = time.time()
    _, state = O.train_step_cpu(w, data, conf)  # warm-up 1 (allocator, thread pool)
    t_one = time.time() - t0
    warm = max(0, min(args.warmup - 1, int(0.25 * budget / max(t_one, 1e-3))))
    for _ in range(warm):
        O.train_step_cpu(w, data, conf, adam_state=state)
    steps = max(1, min(args.steps, int(0.7 * budget / max(t_one, 1e-3))))
    t0 = time.time()
    for _ in range(steps):
        O.train_step_cpu(w, data, conf, adam_state=state)
    dt = (time.time() - t0) / steps
    val = 1.0 / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "steps_requested": args.steps, "warmup": warm + 1, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"LightGlue train step N={N_KPTS} d={D_DESC} L={N_LAYERS} (configs[2]); one pair per step",
                   "pairs_per_step": 1},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "host_cpus": os.cpu_count(),
                         "sample": f"{steps} train steps of 1 pair (N={N_KPTS}, L={N_LAYERS}) after {warm + 1} warm-up"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(seconds=25.0):
    """Bounded CPU sample for the default run's `cpu_baseline` object (rank 0, N=1 only)."""
    from gluefactory_b200 import synthetic
    from oracle import lightglue_oracle as O

    torch.set_num_threads(host_threads())
    conf = dict(synthetic.DEFAULT_CONF)
    w = {k: v.clone().requires_grad_(True) for k, v in synthetic.make_weights(conf, seed=0).items()}
    data = synthetic.make_pairs(1, N_KPTS, seed=1234)
    t0 = time.time()
    _, state = O.train_step_cpu(w, data, conf)
    t_one = time.time() - t0
    steps = max(1, min(4, int(seconds / max(t_one, 1e-3)) - 1))
    t0 = time.time()
    for _ in range(steps):
        O.train_step_cpu(w, data, conf, adam_state=state)
    dt = (time.time() - t0) / steps
    return {"value": 1.0 / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "host_cpus": os.cpu_count(),
            "sample": f"{steps} train steps of 1 pair (N={N_KPTS}, L={N_LAYERS}, fp32, oracle port) after 1 warm-up"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("LGB200_BENCH_BATCH", "32")), help="pairs per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying "
                    "one CUDA graph of the whole training step")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"

    import torch.distributed as dist

    from gluefactory_b200 import _lib, synthetic
    from gluefactory_b200.matchers.lightglue import LightGlue
    from gluefactory_b200.trainer import MatcherTrainer

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    conf = dict(synthetic.DEFAULT_CONF, precision=args.precision)
    model = LightGlue(conf)
    model.load_state_dict(synthetic.make_weights(conf, seed=0), strict=False)
    model = model.to(dev)
    trainer = MatcherTrainer(model, lr=1e-4)

    B = args.batch
    pool = [pin(synthetic.make_pairs(B, N_KPTS, seed=1234 + 1000 * rank + i)) for i in range(2)]
    pool_dev = [synthetic.to_device(p, dev) for p in pool]
    h2d = nbytes(pool[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    use_graph = not args.no_graph
    l0 = _lib.launch_count
    trainer.step(pool_dev[0])  # also the per-step kernel count (the graph replays exactly these launches)
    launches_per_step = _lib.launch_count - l0
    if use_graph:
        trainer.capture(pool_dev[0], dev, warmup=2)

    def step_resident(i):
        if use_graph:
            trainer.step_graphed(pool_dev[i % len(pool_dev)])  # device-to-device copy into the static inputs + replay
        else:
            trainer.step(pool_dev[i % len(pool_dev)])

    host_loss = []

    def step_e2e(i):
        if use_graph:
            # H2D of this step's inputs from pinned memory: queued one step ahead on a copy stream (prefetch=), so the
            # transfer of batch i+1 overlaps the compute of batch i; every step still moves its own h2d bytes
            loss, _ = trainer.step_graphed(pool[i % len(pool)], prefetch=pool[(i + 1) % len(pool)])
        else:
            loss, _ = trainer.step(pool[i % len(pool)], device=dev)
        host_loss.append(loss.item())                             # D2H read of the step's result

    for i in range(args.warmup):
        step_resident(i)
    with ClockSampler(local) as clk:
        ms = timed(step_resident, args.steps)
        launches = launches_per_step * args.steps
    clocks = clk.summary()
    # dominant kernel, timed live with CUDA events on the launching stream (second pass, so the
    # event records do not perturb the headline number)
    roof = None
    if not args.no_kernel_timing:
        roof = kernel_roofline(trainer, pool_dev, B, dev)
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)

    pairs = B * world * args.steps
    value = pairs / (ms / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"LightGlue matcher train step, N=M={N_KPTS} keypoints, d={D_DESC}, L={N_LAYERS}, H={N_HEADS} "
                               "(BASELINE.json configs[2]); forward+loss+backward+all-reduce+Adam",
                   "pairs_per_gpu_per_step": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "launch": "cuda-graph replay of the whole step" if use_graph else "eager (host launches)",
                   "l2": "per-step working set (activations + N x N similarities, >1 GB) exceeds the 126 MB L2; inputs rotate",
                   "attention_roofline_frac": value / world * ATTN_FLOPS_PER_PAIR / (measured_peaks()["tflops_sustained"] * 1e12)},
        "clocks": clocks,
        "e2e": {"value": pairs / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
    }
    if roof is not None:
        line["roofline"] = roof
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down without destroy_process_group(): with NCCL work captured inside a live CUDA graph the
        # communicator teardown was observed to hang on this stack.  Drop the graph, sync, barrier, hard-exit.
        trainer._graph = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def kernel_roofline(trainer, pool_dev, B, dev):
    """Average device time of the dominant kernel (the tcgen05 attention forward, 36 launches per
    step) measured with CUDA events around each launch, vs its algorithmic FLOPs."""
    from gluefactory_b200 import _lib

    name = os.environ.get("LGB200_ROOFLINE_ENTRY", "lgb200_attn_bwd")
    _lib.timed_events.clear()
    _lib.timed_entry = name
    for i in range(2):
        trainer.step(pool_dev[i % len(pool_dev)])
    torch.cuda.synchronize()
    _lib.timed_entry = None
    times = [s.elapsed_time(e) for s, e, _ in _lib.timed_events]
    _lib.timed_events.clear()
    if not times:
        return None
    avg_ms = sum(times) / len(times)
    # one launch = all heads of the [image0; image1] batch: 2B sequences of N tokens, H heads.
    # algorithmic FLOPs per launch (no recompute): fwd 2 GEMMs, bwd 4 GEMMs (5 for self-attention's dK);
    # we count fwd = 4 N^2 d per (seq, head), bwd = 2x fwd.
    per_head = 4 * N_KPTS * N_KPTS * 64
    flops = per_head * N_HEADS * 2 * B * (2 if name == "lgb200_attn_bwd" else 1)
    peaks = measured_peaks()
    achieved = flops / (avg_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            per_seq = json.load(f).get(name + "_per_sequence")
        traffic = per_seq * 2 * B if per_seq else None  # one launch covers the 2B sequences of the batch
    return {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
            "frac": achieved / peaks["tflops_sustained"], "traffic": traffic, "avg_launch_ms": avg_ms,
            "launches_timed": len(times), "peak_source": peaks["source"] + ", sustained (kernel timed inside a long step)"}


if __name__ == "__main__":
    main()
