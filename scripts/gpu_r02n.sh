#!/bin/bash
# round-2 GPU session N: configs[4] with real integrated gradients on one GPU; bench line with the streaming kernels
cd "$(dirname "$0")/.."
O=gpurun_out/r02n; mkdir -p $O
timeout 900 python scripts/cfg5_longform.py --ig 1 > $O/cfg5_ig_1gpu.json 2> $O/cfg5.err; echo "cfg5 rc=$?" | tee -a $O/summary.txt
cut -c1-1500 $O/cfg5_ig_1gpu.json; tail -3 $O/cfg5.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch']); print(json.dumps(d['kernels'])[:1500]); print(d['vocoder'])"
