"""Append the records of gpurun_out/first_touch_log.jsonl (overwritten by every gpurun call) to the committed log
profiles/r2_first_touch_log.jsonl, skipping records already there (tag + pid + uptime)."""
import json
import os

src, dst = "gpurun_out/first_touch_log.jsonl", "profiles/r2_first_touch_log.jsonl"
key = lambda r: (r["tag"], r["pid"], r.get("uptime_s"))   # noqa: E731
have = {key(json.loads(l)) for l in open(dst)} if os.path.exists(dst) else set()
new = [l for l in open(src) if key(json.loads(l)) not in have] if os.path.exists(src) else []
with open(dst, "a") as f:
    f.writelines(new)
recs = [json.loads(l) for l in open(dst)]
print(f"{len(new)} new records; {len(recs)} in total, {sum(1 for r in recs if r.get('events'))} with events")
