#!/bin/bash
# round-2 GPU session ZZO: final state - GPU suite, smoke, bench line, launch list of the same command, reference arm
cd "$(dirname "$0")/.."
O=gpurun_out/r02zzo; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -2 $O/pytest_all.log | head -1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
tail -1 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench.json')); print('value', round(d['value']), 'us/step', round(1000*d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'wave', round(d['e2e_wave_only']['value']), 'roofline', round(d['roofline']['frac'],4), round(d['roofline']['us_per_launch'],2), 'launches', d['gpu_launches']); k=d['kernels']; print({n:(round(k[n]['us'],2), round(k[n]['frac'],3)) for n in ('stft_X','stft_X_mag_phase','istft')}); print('b256', {n:(round(v['us'],1), round(v['frac'],3)) for n,v in k['batch256'].items() if isinstance(v,dict)}); print('b1024', {n:(round(v['us'],1), round(v['frac'],3)) for n,v in k['batch1024'].items() if isinstance(v,dict)}); print('refdef', {n:(round(v['us'],1), round(v['frac'],3)) for n,v in k['reference_default_geometry'].items() if isinstance(v,dict)}); print(k['mel_frontend']['us'], d['vocoder']['clips_per_s'], d['vocoder']['frac'], d['cpu_baseline']['value'], d['clocks'])"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm rc=$?" | tee -a $O/summary.txt
cut -c1-200 $O/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_steps20.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
