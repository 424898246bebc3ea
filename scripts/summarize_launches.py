"""ncu `--metrics gpu__time_duration.sum --csv` launch list -> markdown table of device time per kernel."""
import collections, csv, sys
src, title = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows:
    if r is hdr or len(r) <= iv or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else v
    name = r[ik].split("(")[0][:70]
    agg[name][0] += v
    agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
n = sum(v[1] for v in agg.values())
print(f"# {title}\n\ntotal {tot/1e3:.2f} ms over {n} launches\n")
print("| device time (us) | share | launches | avg (us) | kernel |\n|---:|---:|---:|---:|---|")
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"| {t:.1f} | {100*t/tot:.1f}% | {c} | {t/c:.1f} | `{k}` |")
lgb = sum(v[0] for k, v in agg.items() if "lgb::" in k)
nv = sum(v[0] for k, v in agg.items() if "nvjet" in k or "cutlass" in k or "cublas" in k)
print(f"\nShares: this library's kernels {100*lgb/tot:.0f} %, cuBLAS projections {100*nv/tot:.0f} %, remaining torch glue {100*(tot-lgb-nv)/tot:.0f} %.")
