#!/bin/bash
# usage: tools/grun.sh <timeout_s> '<command>'   -- gpurun with retries while the pod answers busy / transient
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out"
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3\|retry in a few minutes"; then sleep 90; continue; fi
  break
done
