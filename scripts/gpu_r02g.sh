#!/bin/bash
# round-2 GPU session G: explain4 at 120 registers + 128-thread normaliser CTAs (co-residency in the pooled schedule)
cd "$(dirname "$0")/.."
O=gpurun_out/r02g; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_explain4.py tests/test_gpu_end_to_end.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest.log
K="timeout 300 python scripts/kbench.py"
$K explain --tag r120_b64 > $O/kbench.jsonl 2> $O/kbench.err
for i in 1 2; do
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n128_$i.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
ADV_NORM_THREADS=256 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n256_$i.json 2>> $O/bench.err
done
timeout 900 python bench.py --steps 2000 --warmup 5 --no-cpu-baseline > $O/bench_n128_long.json 2>> $O/bench.err
cut -c1-330 $O/kbench.jsonl
for f in $O/bench_*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],5), 'burst', d['run']['burst_us_per_step'], 'explain us', round(d['roofline']['us_per_launch'],2), 'e2e', round(d['e2e']['value']))"; done
