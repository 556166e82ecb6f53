#!/bin/bash
# round-2 GPU session ZQ: single-block fast path of the LMAC reduction; bench line; metric parity tests
cd "$(dirname "$0")/.."
O=gpurun_out/r02zq; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_end_to_end.py tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest.log
for i in 1 2; do
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$i.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench$i.json')); print('value', round(d['value']), 'ms/step', d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch'], 'burst', d['run']['burst_us_per_step'])"
done
