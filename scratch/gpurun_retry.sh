#!/bin/bash
# usage: scratch/gpurun_retry.sh <out-file> <timeout> [--gpus N] -- '<command>'   (retries while the pod answers "busy")
out=$1; shift; to=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$1" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "[retry] finished rc=$rc after $i attempt(s)" >> "$out"; exit $rc; fi
  sleep 90
done
echo "[retry] gave up" >> "$out"
