#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged).  usage: tools/gpurun_retry.sh TIMEOUT 'command' [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then OUT=$(gpurun --timeout $T -- "$CMD" 2>&1); else OUT=$(gpurun --gpus $G --timeout $T -- "$CMD" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient\|no box\|slot free"; then sleep 90; continue; fi
  echo "$OUT" | tail -80; exit 0
done
echo "$OUT" | tail -20
