#!/bin/bash
# Runs on the GPU box (gpurun): the evidence committed under profiles/ for this round.
#   gpurun --timeout 1500 -- 'bash tools/round_artifacts.sh r1b'
tag=${1:-r1b}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; tail -3 $out/${tag}_pytest_gpu.log
python bench.py > $out/${tag}_bench_n1.log 2>&1; grep '^{' $out/${tag}_bench_n1.log > $out/${tag}_bench_n1.json
python bench.py --impl reference > $out/${tag}_bench_reference.log 2>&1; grep '^{' $out/${tag}_bench_reference.log > $out/${tag}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_bench_under_ncu.log 2>&1
for k in expand_mask_kernel matvec_shared_kernel challenge_kernel sign_tail_sparse_kernel resolve_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $out/${tag}_$k python tools/sign_once.py 2 65536 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:drain_kernel -c 1 -f -o $out/${tag}_drain_kernel python tools/e2e_sign_bench.py 2 65536 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ntt_tma_kernel -c 2 -f -o $out/${tag}_ntt python tools/quick_bench.py > /dev/null 2>&1
compute-sanitizer --tool memcheck python tools/kernel_tour.py small 16 > $out/${tag}_sanitizer_memcheck.log 2>&1; tail -2 $out/${tag}_sanitizer_memcheck.log
compute-sanitizer --tool racecheck python tools/kernel_tour.py small 16 > $out/${tag}_sanitizer_racecheck.log 2>&1; tail -2 $out/${tag}_sanitizer_racecheck.log
compute-sanitizer --tool memcheck python tools/kernel_tour.py small > $out/${tag}_sanitizer_memcheck_spec.log 2>&1; tail -1 $out/${tag}_sanitizer_memcheck_spec.log
ls -la $out | tail -30
