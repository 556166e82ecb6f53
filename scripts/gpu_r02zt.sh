#!/bin/bash
# round-2 GPU session ZT: final validation - whole GPU suite, smoke, bench line + launch list of the same command
cd "$(dirname "$0")/.."
O=gpurun_out/r02zt; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench.json')); print('value', round(d['value']), 'ms/step', d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch'], 'launches', d['gpu_launches']); k=d['kernels']; print({n:(round(k[n]['us'],2), round(k[n]['frac'],3)) for n in ('stft_X','stft_X_mag_phase','istft')}); print('b256', {n:(round(v['us'],1), round(v['frac'],3)) for n,v in k['batch256'].items() if isinstance(v,dict)}); print('refdef', {n:(round(v['us'],1), round(v['frac'],3)) for n,v in k['reference_default_geometry'].items() if isinstance(v,dict)}); print(k['mel_frontend']['us'], k['mel_frontend']['two_launch_us']); print(d['vocoder']); print(d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['clocks'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_steps20.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
