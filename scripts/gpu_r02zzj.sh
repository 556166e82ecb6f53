#!/bin/bash
# round-2 GPU session ZZJ: mask gains of the generation-2/3 fused kernels through the .approx.ftz MUFU forms - suite + A/B at a
# geometry they serve (n_fft 1024, hop 256, full hann window), then the final bench line and launch list
cd "$(dirname "$0")/.."
O=gpurun_out/r02zzj; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -2 $O/pytest_all.log | head -1
for v in default oldgains default oldgains; do
  if [ $v = default ]; then unset ADV_LIB_PATH; else export ADV_LIB_PATH=$PWD/xai-audio-deepfakes_b200/libaddvisor_sm100.$v.so; fi
  timeout 300 python scripts/kbench.py explain --nfft 1024 --hop 256 --win hann --winlen 1024 --tag $v 2>/dev/null | tail -1 | cut -c1-300 | tee -a $O/kbench_explain3.jsonl
done
unset ADV_LIB_PATH
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench.json')); print('value', round(d['value']), 'us/step', round(1000*d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],4), round(d['roofline']['us_per_launch'],2)); k=d['kernels']; print({n:(round(k[n]['us'],2), round(k[n]['frac'],3)) for n in ('stft_X','stft_X_mag_phase','istft')}); print('refdef', {n:(round(v['us'],1), round(v['frac'],3)) for n,v in k['reference_default_geometry'].items() if isinstance(v,dict)}); print(k['mel_frontend']['us'], d['vocoder']['clips_per_s'], d['cpu_baseline']['value'], d['clocks'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_steps20.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
