#!/bin/bash
# round-2 GPU session ZW: warp-private mask columns in the streaming explain kernel (no CTA-wide mask rendezvous)
cd "$(dirname "$0")/.."
O=gpurun_out/r02zw; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_explain4.py tests/test_gpu_parity.py -x -q -m gpu > $O/pytest_e4.log 2>&1; echo "pytest e4+parity rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_e4.log
for i in 1 2; do
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$i.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench$i.json')); print('value', round(d['value']), 'ms/step', d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch'], 'burst', d['run']['burst_us_per_step'], 'launches', d['gpu_launches'])"
done
timeout 600 python scripts/stress_e4.py 200 > $O/stress.log 2>&1; echo "stress rc=$?" | tee -a $O/summary.txt
grep -v "^MISMATCH\|per-row\|stats diff" $O/stress.log | tail -4
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_all.log
