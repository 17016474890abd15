mkdir -p gpurun_out /tmp/kat
python - <<'PY'
import sys, os
sys.path.insert(0, "tests")
import oracle_lib as ol
K = ol.kat(2)
for stem in ("rho", "k", "tr", "s1", "s2", "t0"):
    open(f"/tmp/kat/{stem}_2.txt", "w").write(K[stem][0].tobytes().hex().upper() + "\n")
PY
for g in 1 2 4 8; do examples/pool_sign /tmp/kat 2 65536 5 $g; done > gpurun_out/r2c_pool_sign.jsonl 2>&1
cat gpurun_out/r2c_pool_sign.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2c_bench_n8.log 2>&1
tail -c 300 gpurun_out/r2c_bench_n8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 --no-configs > gpurun_out/r2c_bench_n4.log 2>&1
tail -c 300 gpurun_out/r2c_bench_n4.log
