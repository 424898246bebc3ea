"""Tiny driver for ncu: a few launches of the attention kernels at the benchmark shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluefactory_b200 import ops
B, N, H = int(os.environ.get("PB", "8")), 2048, 4
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v, go = (torch.randn(B, N, H, 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(4))
for _ in range(3):
    out, lse = ops.attn_fwd(q, k, v, B // 2, 0.125)
    dq, dk, dv = ops.attn_bwd(q, k, v, out, lse, go, B // 2, 0.125)
torch.cuda.synchronize()
# event timing without the profiler
def t(fn, n=20):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
tf = t(lambda: ops.attn_fwd(q, k, v, B // 2, 0.125))
tb = t(lambda: ops.attn_bwd(q, k, v, out, lse, go, B // 2, 0.125))
fl = 4 * N * N * 64 * B * H
print(f"fwd {tf*1e3:.1f} us  {fl/tf/1e9:.0f} TFLOP/s | bwd {tb*1e3:.1f} us  {2*fl/tb/1e9:.0f} TFLOP/s (algorithmic)")
try:  # per-kernel durations (CUPTI sees the ctypes-launched kernels too)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            ops.attn_fwd(q, k, v, B // 2, 0.125)
            ops.attn_bwd(q, k, v, out, lse, go, B // 2, 0.125)
        torch.cuda.synchronize()
    for ev in prof.key_averages():
        print(f"  {ev.key[:60]:60s} n={ev.count:3d} avg={ev.device_time_total / max(ev.count,1):8.1f} us")
except Exception as ex:  # noqa: BLE001
    print("profiler unavailable:", ex)
