#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [gpurun args...] -- 'command'
# Retries while the pod answers busy (exit code 3 AND nothing charged).  A call that lost its box also ends with
# exit code 3 but WAS charged and counts as a strike: never retry that one blindly (round 2 lost gpurun this way —
# three identical attempts, three strikes).
log="$1"; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  if ! grep -q "charged=0.0s" "$log"; then
    echo "gpurun_retry: rc 3 with box time charged (lost box?) - not retrying" >> "$log"
    exit 4
  fi
  sleep 90
done
exit 3
