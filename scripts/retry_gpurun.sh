#!/bin/bash
# retries a gpurun call while the pod answers busy (exit 3 / "transient")
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -${TAILN:-60}
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  break
done
