#!/bin/bash
# keep asking for a GPU box until a call is actually run (the pod answers "transient" while its slots are busy)
# usage: gpurun_retry.sh <max_tries> <gpurun args...>
tries=$1; shift
for i in $(seq 1 "$tries"); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if ! echo "$out" | grep -q "status=transient"; then echo "$out"; exit 0; fi
  echo "[retry $i] $(echo "$out" | grep -E 'busy|draining|backing off' | head -1)"
  sleep 150
done
echo "gave up after $tries tries"
