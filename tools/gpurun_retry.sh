#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <command...>   — retries while the pod answers "busy" (exit code 3)
# GPURUN_GPUS=N in the environment asks for N GPUs of one box
T=$1; shift
G=${GPURUN_GPUS:-1}
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry $i] pod busy, sleeping 90 s"
  sleep 90
done
exit 3
