#!/bin/bash
# round-2 GPU session ZZR: last check of the committed state - GPU suite, smoke, one bench line
cd "$(dirname "$0")/.."
O=gpurun_out/r02zzr; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -2 $O/pytest_all.log | head -1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
tail -1 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench.json')); print('value', round(d['value']), 'us/step', round(1000*d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],4), round(d['roofline']['us_per_launch'],2), 'launches', d['gpu_launches'], d['clocks'])"
