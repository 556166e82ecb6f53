#!/bin/bash
# round-2 GPU session ZZP: pooled schedule with 1 / 2 / 3 explain streams on the final kernels
cd "$(dirname "$0")/.."
O=gpurun_out/r02zzp; mkdir -p $O
for s in 2 1 3 4 2; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --streams $s > $O/bench_s$s.json 2> $O/bench_s$s.err
  python -c "
import json; d=json.load(open('$O/bench_s$s.json')); print(json.dumps({'streams':$s,'value':round(d['value']),'us_per_step':round(1000*d['ms_per_step'],2),'burst':d['run']['burst_us_per_step']}))" | tee -a $O/streams.jsonl
done
