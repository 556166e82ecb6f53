#!/bin/bash
# round-2 GPU session ZU: normaliser + metric reduction in one launch - parity, bench (value), 1 024-clip kernel lines
cd "$(dirname "$0")/.."
O=gpurun_out/r02zu; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_all.log
for i in 1 2; do
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$i.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench$i.json')); print('value', round(d['value']), 'ms/step', d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch'], 'burst', d['run']['burst_us_per_step'], 'launches', d['gpu_launches']); print('b1024', {n:(round(v['us'],1), round(v['frac'],3)) for n,v in d['kernels']['batch1024'].items() if isinstance(v,dict)} if 'error' not in d['kernels']['batch1024'] else d['kernels']['batch1024'])"
done
tail -3 $O/bench.err
