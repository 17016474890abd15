#!/bin/bash
# Runs tools/e2e_sign_bench.py on N GPUs at once (one process per GPU) to see how the host path scales
# when every GPU streams signatures to host memory at the same time:  bash tools/e2e_multi.sh N [env...]
n=$1; shift
for i in $(seq 0 $((n-1))); do
  env CUDA_VISIBLE_DEVICES=$i "$@" python tools/e2e_sign_bench.py 2 65536 40 > /tmp/e2e_$i.log 2>&1 &
done
wait
for i in $(seq 0 $((n-1))); do tail -1 /tmp/e2e_$i.log | sed "s/^/gpu$i /" | cut -c1-330; done
