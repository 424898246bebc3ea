"""Reads the clock64 pipeline trace of CTA 0 of an attention kernel and prints the per-tile time stamps.
Build with LGB200_EXTRA_NVCC_FLAGS=-DLGB_TRACE=1 (dKV backward kernel) or -DLGB_TRACE=3 (forward kernel)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluefactory_b200 import ops, _lib
B, N, H = 8, 2048, 4
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v, go = (torch.randn(B, N, H, 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(4))
for _ in range(2):
    out, lse = ops.attn_fwd(q, k, v, B // 2, 0.125)
    ops.attn_bwd(q, k, v, out, lse, go, B // 2, 0.125)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * 1024)()
rc = lib.lgb200_debug_read_trace(buf, 1024)
t = torch.tensor(list(buf)).view(4, 64, 4)
t0 = int(t[t > 0].min())
fwd = len(sys.argv) > 1 and sys.argv[1] == "fwd"
names = ["producer", "mma: P-wait done, PV issued, S(j+2) issued", "softmax w0: S-wait done, ld done, exps done, arrive done",
         "softmax w3: same"] if fwd else ["producer: empty-wait done",
         "mma: pds-wait done, dV/dK issued, (dQ issued), sp(i+2) issued",
         "softmax w0: sp-wait done, math done, arrive done, drain done",
         "softmax w15: sp-wait done, math done, arrive done, drain done"]
for role in range(4):
    print(names[role])
    for i in range(32):
        row = [int(x) - t0 if x > 0 else -1 for x in t[role, i]]
        print(f"  tile {i:2d}: " + " ".join(f"{x:7d}" for x in row))
for r, name in ((40, "CTA 0"), (41, "CTA 300")):
    c0, c1, g0, g1 = (int(x) for x in t[0, r])
    print(f"{name}: life {c1 - c0} clk = {g1 - g0} ns  ({(c1 - c0) / max(g1 - g0, 1):.3f} GHz); entry at {c0 - t0} clk rel. to first stamp")
