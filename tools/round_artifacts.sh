#!/bin/bash
# Runs on the GPU box (gpurun): the evidence committed under profiles/ for this round.
#   gpurun --timeout 1700 -- 'bash tools/round_artifacts.sh r2d [tests|ncu]'
# ncu reports are summarised ON THE BOX (tools/ncu_summary.py) and deleted: gpurun brings back at most 64 MiB.
tag=${1:-r2d}
what=${2:-all}
out=gpurun_out
mkdir -p $out
set -x
if [ "$what" = all ] || [ "$what" = tests ]; then
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; tail -3 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench_n1.log 2>&1; grep '^{' $out/${tag}_bench_n1.log > $out/${tag}_bench_n1.json
python bench.py --impl reference > $out/${tag}_bench_reference.log 2>&1; grep '^{' $out/${tag}_bench_reference.log > $out/${tag}_bench_reference.json
compute-sanitizer --tool memcheck python tools/kernel_tour.py small 16 > $out/${tag}_sanitizer_memcheck.log 2>&1; tail -2 $out/${tag}_sanitizer_memcheck.log
compute-sanitizer --tool racecheck python tools/kernel_tour.py small 16 > $out/${tag}_sanitizer_racecheck.log 2>&1; tail -2 $out/${tag}_sanitizer_racecheck.log
grep -E "Race reported|hazard" $out/${tag}_sanitizer_racecheck.log | sed 's/.*in kernel//' | sort | uniq -c | sort -rn | head -20
compute-sanitizer --tool memcheck python tools/kernel_tour.py small > $out/${tag}_sanitizer_memcheck_spec.log 2>&1; tail -1 $out/${tag}_sanitizer_memcheck_spec.log
compute-sanitizer --tool racecheck python tools/kernel_tour.py small 16 fused > $out/${tag}_sanitizer_racecheck_fused.log 2>&1; tail -1 $out/${tag}_sanitizer_racecheck_fused.log
head -c 200000 $out/${tag}_sanitizer_racecheck.log > $out/${tag}_sanitizer_racecheck.head; mv $out/${tag}_sanitizer_racecheck.head $out/${tag}_sanitizer_racecheck.log
head -c 100000 $out/${tag}_sanitizer_racecheck_fused.log > $out/${tag}_sanitizer_racecheck_fused.head; mv $out/${tag}_sanitizer_racecheck_fused.head $out/${tag}_sanitizer_racecheck_fused.log
fi
if [ "$what" = all ] || [ "$what" = ncu ]; then
# launch list of the bench command itself (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 1 > $out/${tag}_bench_under_ncu.log 2>&1
cap() {   # cap <name> <kernel regex> <launches> <units> <unit name> <command...>
  local name=$1 rx=$2 cnt=$3 units=$4 uname=$5; shift 5
  ncu --set full --clock-control none --import-source on -k regex:$rx -c $cnt -f -o /tmp/${tag}_$name "$@" > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/${tag}_$name.ncu-rep $out/${tag}_$name --units-per-launch $units --unit-name $uname > /dev/null 2>&1
  ncu -i /tmp/${tag}_$name.ncu-rep --page source --csv 2>/dev/null | head -400 > $out/${tag}_${name}_source_head.csv
  rm -f /tmp/${tag}_$name.ncu-rep
}
for k in expand_mask_kernel matvec_shared_kernel challenge_kernel sign_tail_sparse_kernel resolve_kernel; do
  cap sign_$k $k 1 65536 slot python tools/sign_once.py 2 65536
done
cap sign_sign_tail_sparse_kernel_L3 sign_tail_sparse_kernel 1 65536 slot python tools/sign_once.py 3 65536
cap sign_drain_kernel drain_kernel 1 15500 signature python tools/e2e_sign_bench.py 2 65536 1
cap ntt 'ntt_' 2 262144 poly python tools/quick_bench.py
cap verify_matvec_item_kernel matvec_item_kernel 1 131072 item python tools/verify_multi_once.py 5 131072
cap keccak_rate_kernel keccak_rate_kernel 1 56832 thread python -c "
import dilithium_b200 as d
print(d.Engine(0).keccak_rate(ctas_per_sm=3, perms=1500, repeats=0))"
fi
ls -la $out | tail -40; du -sh $out
