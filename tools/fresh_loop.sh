#!/bin/bash
# Fresh-process repro loop for the rare wrong first evaluation: TOTAL new processes, PAR at a time, each
# builds a fresh engine and compares its FIRST evaluation bitwise with the outputs stored by the first process.
#   bash tools/fresh_loop.sh TAG TOTAL PAR
tag=${1:-loop}; total=${2:-200}; par=${3:-8}
mkdir -p gpurun_out
ref=gpurun_out/first_touch_${tag}_ref.npz
python tools/first_touch.py ${tag}_ref --save $ref --quiet || echo "reference process reported an event"
fails=0; done_n=0
t0=$(date +%s)
while [ $done_n -lt $total ]; do
  pids=()
  for k in $(seq 1 $par); do
    [ $((done_n + k)) -gt $total ] && break
    timeout 120 python tools/first_touch.py ${tag}_$((done_n + k)) --expect $ref --quiet &
    pids+=($!)
  done
  for p in "${pids[@]}"; do wait $p || fails=$((fails + 1)); done
  done_n=$((done_n + ${#pids[@]}))
done
echo "[fresh_loop $tag] $done_n fresh processes ($par at a time), $fails with an event, $(( $(date +%s) - t0 )) s"
