#!/bin/bash
# usage: profiles/grun.sh <script> <log> [gpus]  -- retries while the pod answers busy / transient
for i in 1 2 3 4 5 6 7 8; do
  if [ -n "$3" ]; then gpurun --gpus $3 --timeout 2400 -- "bash $1" > $2 2>&1; else gpurun --timeout 2400 -- "bash $1" > $2 2>&1; fi
  if grep -qE "status=(ok|fail)" $2; then break; fi
  sleep 90
done
