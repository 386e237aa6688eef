#!/bin/bash
# retries a gpurun call while the pod answers "busy" (exit 3); usage: gpu_retry.sh <timeout> <script>
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $1 -- "bash $2" > /tmp/gpu_retry_last.txt 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then cat /tmp/gpu_retry_last.txt | tail -80; exit $rc; fi
  sleep 150
done
echo "gave up: pod busy"; exit 3
