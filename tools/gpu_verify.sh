#!/bin/bash
# One gpurun call = one box acquisition (~25 s charged before anything runs): bundle what a verification needs.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_verify.sh [tag]'
# Writes gpurun_out/<tag>_*.{json,csv,log}; prints the tails.  Every step runs under its own `timeout`.
tag=${1:-check}
mkdir -p gpurun_out
timeout 200 python tools/first_touch.py ${tag}_first 2>&1 | tail -1 | cut -c1-200
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench.err
python - "$tag" <<'PY'
import json, sys
d = json.loads([l for l in open(f"gpurun_out/{sys.argv[1]}_bench_cfg2.json") if l.startswith("{")][-1])
r = d["roofline"]
print(f"bench: {d['value'] / 1e6:.3f} M atom-steps/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value'] / 1e6:.3f} M  "
      f"GEMM {r['gemm_ms_per_step']:.3f} ms ({r['frac']:.3f} of peak)  phases {r['phase_ms']}  clocks {d['clocks']}")
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${tag}_ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/${tag}_launches_cfg2.csv 2>/dev/null | head -12
# back in the build container: python tools/merge_first_touch.py appends gpurun_out/first_touch_log.jsonl to profiles/
