#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>' [gpus]
# Retries `gpurun` while the pod answers "busy / no box" (exit code 3, nothing charged).
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $i answered busy; sleeping 90 s"; sleep 90
done
exit 3
