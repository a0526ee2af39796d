"""Launch shares of ONE steady-state training step from an `ncu --metrics gpu__time_duration.sum --csv` launch list of
bench.py: the launches between the last two k_adam_multi* launches (= the last timed step).
    python tools/launch_shares.py launches.csv [out.txt]"""
import collections
import csv
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v, u = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
    seq.append((row["Kernel Name"].split("(")[0].replace("void ", "").replace("lfs::", ""), v))
adam = [i for i, (k, _) in enumerate(seq) if k.startswith("k_adam_multi")]
assert len(adam) >= 2, "need two optimiser steps in the capture"
step = seq[adam[-2] + 1:adam[-1] + 1]
agg, cnt = collections.OrderedDict(), collections.Counter()
for k, v in step:
    agg[k] = agg.get(k, 0.0) + v
    cnt[k] += 1
tot = sum(agg.values())
views = max(1, cnt.get("k_preprocess_fwd", 1))
out = [f"One steady-state step (launches between the last two k_adam_multi): {len(step)} launches, {views} views, "
       f"{tot / 1e3:.3f} ms of kernel time = {tot / 1e3 / views:.3f} ms per view.",
       "Times are per-launch ncu durations (cold caches, serialised, --clock-control none): the SHARES are what must agree with",
       "bench.py's stage_ms_per_view, not the absolute values.", ""]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    out.append(f"{100 * v / tot:6.2f} %  {v / 1e3 / views:8.4f} ms/view  {cnt[k]:4d} x  {k}")
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
