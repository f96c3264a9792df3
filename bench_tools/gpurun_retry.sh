#!/bin/bash
# usage: gpurun_retry.sh <outfile> <gpurun args...>   — retries while the pod is busy (exit code 3)
out="$1"; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "[retry] done rc=$rc" >> "$out"; exit $rc; fi
  sleep 45
done
echo "[retry] gave up" >> "$out"; exit 3
