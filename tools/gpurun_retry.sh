#!/bin/bash
# usage: [GPUS=N] gpurun_retry.sh <logfile> <timeout> <command...>   -- retries while the pod is busy (nothing is charged then)
log="$1"; shift; to="$1"; shift
extra=""
if [ -n "$GPUS" ]; then extra="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" $extra -- "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log"; then sleep 45; continue; fi
  break
done
echo "DONE after $i tries" >> "$log"
