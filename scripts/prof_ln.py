"""Driver for ncu / event timing of the LayerNorm+GELU kernels at the benchmark shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluefactory_b200 import ops
T, W = int(os.environ.get("PT", "131072")), 512
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(T, W, device="cuda", generator=g).to(torch.bfloat16)
dy = torch.randn(T, W, device="cuda", generator=g).to(torch.bfloat16)
gamma = torch.randn(W, device="cuda", generator=g) * 0.1 + 1
beta = torch.randn(W, device="cuda", generator=g) * 0.1
def t(fn, n=10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3
y, mean, rstd = ops.ln_gelu_fwd(x, gamma, beta, 1e-5)
tf = t(lambda: ops.ln_gelu_fwd(x, gamma, beta, 1e-5))
tb = t(lambda: ops.ln_gelu_bwd(dy, x, gamma, beta, mean, rstd, want_dxsum=True))
print(f"ln_gelu fwd {tf:.1f} us ({T*W*4/tf/1e6:.2f} TB/s)  bwd {tb:.1f} us ({T*W*6/tb/1e6:.2f} TB/s)")
