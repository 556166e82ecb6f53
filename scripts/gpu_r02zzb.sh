#!/bin/bash
# round-2 GPU session ZZB: final validation - whole GPU suite, smoke, bench line + launch list of the same command, the
# reference arm, and a fresh --set full capture of the production kernels (1 024 clips) for profiles/traffic.json
cd "$(dirname "$0")/.."
O=gpurun_out/${OUT:-r02zzb}; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
tail -2 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench.json')); print('value', round(d['value']), 'us/step', round(1000*d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['run'].get('e2e_host_affinity'), d['pcie'], 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch'], 'launches', d['gpu_launches']); k=d['kernels']; print({n:(round(k[n]['us'],2), round(k[n]['frac'],3)) for n in ('stft_X','stft_X_mag_phase','istft')}); print('refdef', {n:(round(v['us'],1), round(v['frac'],3)) for n,v in k['reference_default_geometry'].items() if isinstance(v,dict)}); print(d['vocoder']); print(d['cpu_baseline'], d['clocks'])"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm rc=$?" | tee -a $O/summary.txt
cut -c1-400 $O/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_steps20.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"explain4_kernel|stft3_kernel|istft4_kernel|mel_fused_kernel" -f -o $O/prof_full python scripts/prof_traffic.py > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $O/summary.txt
tail -2 $O/ncu_full.log
ncu -i $O/prof_full.ncu-rep --page raw --csv > $O/raw_all.csv 2>/dev/null
ncu -i $O/prof_full.ncu-rep --page source --csv --kernel-name regex:explain4 > $O/e4_source.csv 2>/dev/null; rm -f $O/prof_full.ncu-rep
ls -la $O
