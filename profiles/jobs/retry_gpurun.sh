#!/bin/bash
# usage: retry_gpurun.sh <timeout> <script> [gpus]   — re-submits while the pod answers "transient" (busy), up to 12 times
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 12); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient"; then echo "[retry $i] busy"; sleep 120; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up"; exit 3
