#!/bin/bash
# Final validation on the GPU box: full GPU test suite, smoke(), sanitizer tour, one bench line.
tag=${1:-final}
out=gpurun_out
mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; tail -2 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
compute-sanitizer --tool memcheck python tools/kernel_tour.py small 16 > $out/${tag}_sanitizer_memcheck.log 2>&1; tail -2 $out/${tag}_sanitizer_memcheck.log
compute-sanitizer --tool racecheck python tools/kernel_tour.py small 16 > $out/${tag}_sanitizer_racecheck.log 2>&1; tail -2 $out/${tag}_sanitizer_racecheck.log
python bench.py > $out/${tag}_bench_n1.log 2>&1; grep '^{' $out/${tag}_bench_n1.log > $out/${tag}_bench_n1.json
python -c "
import json; d=json.load(open('$out/${tag}_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['step_profile_ms'])"
