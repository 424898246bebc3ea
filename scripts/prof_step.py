"""Per-kernel device time of one eager training step (torch.profiler / CUPTI, warm caches), sorted by total."""
import os, sys, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluefactory_b200 import synthetic
from gluefactory_b200.matchers.lightglue import LightGlue
from gluefactory_b200.trainer import MatcherTrainer
from torch.profiler import profile, ProfilerActivity
B = int(os.environ.get("PB", "32"))
dev = torch.device("cuda", 0)
conf = dict(synthetic.DEFAULT_CONF, precision="bf16")
model = LightGlue(conf)
model.load_state_dict(synthetic.make_weights(conf, seed=0), strict=False)
from gluefactory_b200.matchers.homography_matcher import HomographyMatcher
trainer = MatcherTrainer(model.to(dev), lr=1e-4, ground_truth=HomographyMatcher({"th_positive": 3.0, "th_negative": 3.0, "transposed_assignment": True}))
data = synthetic.to_device(synthetic.make_pairs(B, 2048, seed=1, with_gt=False), dev)
for _ in range(3):
    trainer.step(data)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    trainer.step(data)
    torch.cuda.synchronize()
rows = [(ev.device_time_total, ev.count, ev.key) for ev in prof.key_averages() if ev.device_time_total > 0]
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"B={B} pairs: total device time {tot/1e3:.2f} ms over {sum(r[1] for r in rows)} launches")
print("| device time (us) | share | launches | avg (us) | kernel |\n|---:|---:|---:|---:|---|")
for t, n, k in rows[:60]:
    print(f"| {t:.0f} | {100*t/tot:.1f}% | {n} | {t/n:.1f} | `{k[:90]}` |")
