"""Which host call sites launch the leftover torch kernels (copies, casts, adds, fills) of the training step?
CPU+CUDA profile of one eager step; for every ATen op whose kernels are not ours: kernel time, count, input shapes
and the innermost repo frames."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from gluefactory_b200 import synthetic  # noqa: E402
from gluefactory_b200.matchers.homography_matcher import HomographyMatcher  # noqa: E402
from gluefactory_b200.matchers.lightglue import LightGlue  # noqa: E402
from gluefactory_b200.trainer import MatcherTrainer  # noqa: E402

B = int(os.environ.get("PB", "32"))
dev = torch.device("cuda", 0)
conf = dict(synthetic.DEFAULT_CONF, precision="bf16")
model = LightGlue(conf)
model.load_state_dict(synthetic.make_weights(conf, seed=0), strict=False)
gt = HomographyMatcher({"th_positive": 3.0, "th_negative": 3.0, "transposed_assignment": True})
trainer = MatcherTrainer(model.to(dev), lr=1e-4, ground_truth=gt)
data = synthetic.to_device(synthetic.make_pairs(B, 2048, seed=1, with_gt=False), dev)
for _ in range(2):
    trainer.step(data)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    trainer.step(data)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    ks = [k for k in getattr(ev, "kernels", []) if "lgb::" not in k.name]
    if not ks or not ev.name.startswith("aten::"):
        continue
    st = [s for s in (ev.stack or []) if ".py" in s and "/torch/" not in s][:3]
    key = (ev.name, str(ev.input_shapes)[:70], " <- ".join(s.split("/")[-1][:48] for s in st), ks[0].name[:40])
    agg[key][0] += sum(k.duration for k in ks)
    agg[key][1] += 1
rows = sorted(((v[0], v[1], k) for k, v in agg.items()), reverse=True)
print(f"torch-launched kernels: {sum(r[0] for r in rows):.0f} us total")
for t, n, (name, shapes, st, kn) in rows[:40]:
    print(f"{t:7.0f} us x{n:3d} {name:22s} {shapes:70s} {kn:40s} {st}")
