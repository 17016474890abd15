#!/bin/bash
# Runs on the GPU box (gpurun): the evidence committed under profiles/ for this round.
#   gpurun --timeout 1700 -- 'bash tools/round_artifacts.sh r2d'
tag=${1:-r2d}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; tail -3 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench_n1.log 2>&1; grep '^{' $out/${tag}_bench_n1.log > $out/${tag}_bench_n1.json
python bench.py --impl reference > $out/${tag}_bench_reference.log 2>&1; grep '^{' $out/${tag}_bench_reference.log > $out/${tag}_bench_reference.json
# launch list of the bench command itself (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 1 > $out/${tag}_bench_under_ncu.log 2>&1
for k in expand_mask_kernel matvec_shared_kernel challenge_kernel sign_tail_sparse_kernel resolve_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $out/${tag}_sign_$k python tools/sign_once.py 2 65536 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:sign_tail_kernel -c 1 -f -o $out/${tag}_sign_sign_tail_kernel_L3 python tools/sign_once.py 3 65536 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:drain_kernel -c 1 -f -o $out/${tag}_sign_drain_kernel python tools/e2e_sign_bench.py 2 65536 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ntt_ -c 2 -f -o $out/${tag}_ntt python tools/quick_bench.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:matvec_item_kernel -c 1 -f -o $out/${tag}_verify_matvec_item_kernel python tools/verify_multi_once.py 5 131072 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:keccak_rate_kernel -c 1 -f -o $out/${tag}_keccak_rate_kernel python -c "
import dilithium_b200 as d
print(d.Engine(0).keccak_rate(ctas_per_sm=3, perms=1500, repeats=0))" > /dev/null 2>&1
compute-sanitizer --tool memcheck python tools/kernel_tour.py small 16 > $out/${tag}_sanitizer_memcheck.log 2>&1; tail -2 $out/${tag}_sanitizer_memcheck.log
compute-sanitizer --tool racecheck python tools/kernel_tour.py small 16 > $out/${tag}_sanitizer_racecheck.log 2>&1; tail -2 $out/${tag}_sanitizer_racecheck.log
compute-sanitizer --tool memcheck python tools/kernel_tour.py small > $out/${tag}_sanitizer_memcheck_spec.log 2>&1; tail -1 $out/${tag}_sanitizer_memcheck_spec.log
ls -la $out | tail -30
