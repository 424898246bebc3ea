"""`.ncu-rep` of the attention-backward kernels inside one eager training step (scripts/gpu_final_r2.sh) ->
profiles/r02_roofline_traffic.json: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per lgb200_attn_bwd call
= one attn_bwd_prep_fused + one attn_bwd_fused (+ one dq_convert when captured), averaged over the captured launches.
Usage: python scripts/summarize_traffic.py gpurun_out/r02_attn_bwd.ncu-rep <sequences_per_launch> [keypoints]"""
import collections
import csv
import io
import json
import subprocess
import sys

rep, seqs = sys.argv[1], int(sys.argv[2])
kpts = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                      "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"],
                     capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Kernel Name" in r)
units = rows[rows.index(hdr) + 1]
ik = hdr.index("Kernel Name")
cols = {m: hdr.index(m) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
         "nsecond": 1e-3}
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for r in rows[rows.index(hdr) + 2:]:
    if len(r) <= max(cols.values()):
        continue
    name = r[ik].split("(")[0].replace("lgb::", "")
    v = {m: float(r[c].replace(",", "")) * scale.get(units[c], 1) for m, c in cols.items()}
    a = agg[name]
    a[0] += v["dram__bytes_read.sum"]
    a[1] += v["dram__bytes_write.sum"]
    a[2] += v["gpu__time_duration.sum"]
    a[3] += 1
per_call, detail = 0.0, {}
for name, (rd, wr, us, n) in agg.items():
    detail[name] = {"launches": n, "dram_read_MB": rd / n / 1e6, "dram_write_MB": wr / n / 1e6, "time_us": us / n}
    per_call += (rd + wr) / n
res = {"_comment": "ncu --set full --clock-control none of the attention-backward kernels inside one eager training step "
                   "(scripts/gpu_final_r2.sh; cold caches, serialised launches); per lgb200_attn_bwd call = the sum over the "
                   "kernels below, each averaged over its captured launches",
       "sequences_per_launch": seqs, "keypoints": kpts, "dram_bytes_per_launch": {"lgb200_attn_bwd": per_call}, "kernels": detail}
json.dump(res, open("profiles/r02_roofline_traffic.json", "w"), indent=1)
print(json.dumps(res, indent=1))
