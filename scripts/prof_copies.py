"""Which host call sites launch the leftover torch copy / add kernels of the training step?  (CPU+CUDA profile with stacks)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluefactory_b200 import synthetic
from gluefactory_b200.matchers.lightglue import LightGlue
from gluefactory_b200.matchers.homography_matcher import HomographyMatcher
from gluefactory_b200.trainer import MatcherTrainer
from torch.profiler import profile, ProfilerActivity
B = int(os.environ.get("PB", "32"))
dev = torch.device("cuda", 0)
conf = dict(synthetic.DEFAULT_CONF, precision="bf16")
model = LightGlue(conf)
model.load_state_dict(synthetic.make_weights(conf, seed=0), strict=False)
trainer = MatcherTrainer(model.to(dev), lr=1e-4, ground_truth=HomographyMatcher({"th_positive": 3.0, "th_negative": 3.0}))
data = synthetic.to_device(synthetic.make_pairs(B, 2048, seed=1, with_gt=False), dev)
for _ in range(2):
    trainer.step(data)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
    trainer.step(data)
    torch.cuda.synchronize()
want = ("aten::copy_", "aten::add", "aten::add_", "aten::clone", "aten::contiguous", "aten::to", "aten::_to_copy", "aten::cat",
        "aten::transpose", "aten::sum", "aten::fill_", "aten::zero_")
rows = []
for ev in prof.key_averages(group_by_stack_n=6):
    if ev.key in want and ev.device_time_total > 100:
        st = [s for s in ev.stack if ".py" in s and "torch/" not in s] or list(ev.stack)
        rows.append((ev.device_time_total, ev.count, ev.key, st[:4]))
rows.sort(reverse=True)
for t, n, k, st in rows[:25]:
    print(f"{t:8.0f} us  x{n:3d}  {k:18s} {' <- '.join(s.split('/')[-1][:60] for s in st)}")
