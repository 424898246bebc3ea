"""Time the persistent Sinkhorn kernels (SuperGlue head): forward (50 iterations) and the reverse
sweep at N = M = 2048 for a few batch sizes; prints us per call and the streaming rate the iterations would need if
every half-iteration read Z once from memory (2 * iters * 4 B * M * N per pair forward, x1.5 for the backward's
dZ read-modify-write)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluefactory_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
iters = 50
for B, M, N in [(1, 2048, 2048), (4, 2048, 2048), (32, 2048, 2048), (8, 512, 512)]:
    sim = torch.randn(B, M, N, device=dev) * 3
    w = torch.randn(B, M + 1, N + 1, device=dev)
    out, uh, vh = ops._log_optimal_transport_fwd(sim, 0.9, iters, keep_potentials=True)
    ops._log_optimal_transport_bwd(sim, 0.9, iters, w, uh, vh)
    torch.cuda.synchronize()
    reps = 5 if B < 32 else 2
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(reps):
        ops._log_optimal_transport_fwd(sim, 0.9, iters, keep_potentials=True)
    ev[1].record()
    for _ in range(reps):
        ops._log_optimal_transport_bwd(sim, 0.9, iters, w, uh, vh)
    ev[2].record()
    torch.cuda.synchronize()
    tf, tb = ev[0].elapsed_time(ev[1]) / reps * 1e3, ev[1].elapsed_time(ev[2]) / reps * 1e3
    pass_bytes = 4.0 * B * M * N
    print(f"B={B} M={M} N={N}: forward {tf:.0f} us ({2 * iters * pass_bytes / tf / 1e3:.0f} GB/s equivalent), "
          f"backward {tb:.0f} us ({2 * iters * pass_bytes * 1.5 / tb / 1e3:.0f} GB/s equivalent)")
