"""Kernel shares of one bench step from an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python tools/launch_shares.py profiles/r1_v19_launches_cfg2.csv
One step = the launches between the last two `embed_kernel` launches (the first kernel of an evaluation)."""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for x in csv.DictReader(lines):
    if x.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(x["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(x["Metric Unit"], v)
    rows.append((re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", ""), v))
starts = [i for i, (k, _) in enumerate(rows) if "embed" in k]
if len(starts) < 2:
    sys.exit("need at least two evaluations in the list")
step = rows[starts[-2]:starts[-1]]
total = sum(v for _, v in step)
agg = collections.OrderedDict()
for k, v in step:
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += v
print(f"{len(step)} launches, {total:.0f} us summed (cold-cache, serialised: read the shares)")
try:
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:64]:64s} {n:3d} x {v / n:8.1f} us  {100 * v / total:5.1f} %")
except BrokenPipeError:   # piped into `head`
    sys.stderr.close()
