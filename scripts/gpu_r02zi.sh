#!/bin/bash
# round-2 GPU session ZI: per-layer epilogue autotune in the vocoder; bench.py line
cd "$(dirname "$0")/.."
O=gpurun_out/r02zi; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -x -q -m gpu > $O/pytest_tc.log 2>&1; echo "pytest tc rc=$?" | tee -a $O/summary.txt
tail -4 $O/pytest_tc.log
for e in auto tma direct; do
  timeout 600 python scripts/bench_vocoder.py --batch 256 --iters 5 --epilogue $e > $O/voc_$e.json 2>> $O/voc.err; cut -c1-330 $O/voc_$e.json
done
timeout 600 python scripts/bench_vocoder.py --batch 64 --iters 5 > $O/voc64.json 2>> $O/voc.err; cut -c1-330 $O/voc64.json
tail -3 $O/voc.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch']); print(d['vocoder'])"
