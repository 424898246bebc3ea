"""Event timing of the device GT-label kernels vs the torch restatement at the benchmark size."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluefactory_b200 import ops, synthetic
B, N = int(os.environ.get("PB", "8")), 2048
d = synthetic.to_device(synthetic.make_pairs(B, N, seed=3, with_gt=False), torch.device("cuda"))
kp0, kp1, H = d["keypoints0"], d["keypoints1"], d["H_0to1"]
def t(fn, n=5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3
tk = t(lambda: ops.gt_matches_from_homography(kp0, kp1, H, 3.0, 3.0))
ts = t(lambda: ops.gt_matches_from_homography(kp0, kp1, H, 3.0, 3.0, dense=False))
tt = t(lambda: synthetic.gt_matches_from_homography(kp0, kp1, H, 3.0, 3.0))
print(f"GT labels, {B} pairs N={N}: kernels {tk:.0f} us (sparse outputs only {ts:.0f} us) | torch restatement {tt:.0f} us")
