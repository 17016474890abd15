#!/bin/bash
# Runs on an N-GPU box (gpurun --gpus N): multi-GPU evidence committed under profiles/.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/multi_gpu_artifacts.sh r2e 8'
tag=${1:-r2e}
G=${2:-8}
mkdir -p gpurun_out /tmp/kat
python - <<'PY'
import sys, os
sys.path.insert(0, "tests")
import oracle_lib as ol
K = ol.kat(2)
for stem in ("rho", "k", "tr", "s1", "s2", "t0"):
    open(f"/tmp/kat/{stem}_2.txt", "w").write(K[stem][0].tobytes().hex().upper() + "\n")
PY
# one process, dil_pool_* over g GPUs, 1 / 2 / 4 batches in flight
for g in ${POOL_GPUS:-1 2 4 8}; do
  [ $g -le $G ] || continue
  for t in ${POOL_IN_FLIGHT:-1 2 4}; do examples/pool_sign /tmp/kat 2 65536 8 $g $t; done
done > gpurun_out/${tag}_pool_sign.jsonl 2>&1
cat gpurun_out/${tag}_pool_sign.jsonl
# one process per GPU (the driver's launch line)
for n in 8 4 2; do
  [ $n -le $G ] || continue
  extra="--no-configs"; [ $n = $G ] && extra=""
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 3 $extra > gpurun_out/${tag}_bench_n$n.log 2>&1
  grep '^{' gpurun_out/${tag}_bench_n$n.log > gpurun_out/${tag}_bench_n$n.json
  python - <<PY
import json
j = json.loads(open("gpurun_out/${tag}_bench_n$n.json").read().strip().splitlines()[-1])
print("N=$n value %.2f M one-at-a-time %.2f M e2e %.2f M (serial %.2f M) in flight %d" % (j["value"] / 1e6, j["one_batch_at_a_time"]["value"] / 1e6,
      j["e2e"]["value"] / 1e6, j["e2e"]["one_call_at_a_time"] / 1e6, j["config"]["batches_in_flight"]), j["clocks"])
PY
done
