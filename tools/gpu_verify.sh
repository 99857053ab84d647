#!/bin/bash
# One gpurun call = one box acquisition (~25 s charged before anything runs): bundle what a verification needs.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_verify.sh [tag] [unverified]'
# Writes gpurun_out/<tag>_*.{json,csv,log}; prints the tails.  "unverified" also runs the tests parked under the
# gpu_unverified marker (tests/test_gpu_unverified.py) first.
tag=${1:-check}
mkdir -p gpurun_out
if [ "$2" = "unverified" ]; then
  # kernels that have run before (seams around verified launchers, backend 2 with many tiles, invariances) ...
  timeout 200 python -m pytest tests -m gpu_unverified -q -s -k "not pipelined" 2>&1 \
    | grep -E "\[seam\]|\[invariance\]|passed|failed|Error|assert" | cut -c1-300
  # ... and, last and under a short leash, the GEMM kernel that has never executed (a deadlock must not eat the budget)
  timeout 90 python -m pytest tests -m gpu_unverified -q -x -k "pipelined" 2>&1 | grep -E "passed|failed|Error|assert" | cut -c1-300
fi
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench.err
python - "$tag" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench_cfg2.json"))
r = d["roofline"]
print(f"bench: {d['value'] / 1e6:.3f} M atom-steps/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value'] / 1e6:.3f} M  "
      f"GEMM {r['gemm_ms_per_step']:.3f} ms ({r['frac']:.3f} of peak)  phases {r['phase_ms']}  clocks {d['clocks']}")
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
wc -l gpurun_out/${tag}_launches_cfg2.csv
