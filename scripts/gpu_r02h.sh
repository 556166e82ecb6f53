#!/bin/bash
# round-2 GPU session H: defaults after the register / normaliser A/B, single-graph short runs, wave-only e2e leg
cd "$(dirname "$0")/.."
O=gpurun_out/r02h; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py --steps 2000 --warmup 5 --no-cpu-baseline > $O/bench_long.json 2>> $O/bench.err
K="timeout 300 python scripts/kbench.py"
$K stft stft3 --nfft 1024 --hop 322 --n 80000 --pool 8 --tag refdef_stft_gen2route > $O/kbench.jsonl 2> $O/kbench.err
cut -c1-400 $O/kbench.jsonl
for f in $O/bench.json $O/bench_long.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],5), 'burst', d['run']['burst_us_per_step'], 'explain us', round(d['roofline']['us_per_launch'],2), 'e2e', round(d['e2e']['value']), 'wave-only', d['e2e_wave_only'], 'pcie', d['pcie'])"; done
tail -3 $O/bench.err
