#!/bin/bash
# round-2 GPU session ZZ: warps without work leave the streaming / warp-autonomous loops (explain4/5, istft4/5, stft3/5),
# tail-empty wait behind the first inverse transform: full suite, race hunts, bench
cd "$(dirname "$0")/.."
O=gpurun_out/r02zz; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_all.log
timeout 600 python scripts/stress_e4.py 200 > $O/stress_e4.log 2>&1; echo "stress e4 rc=$?" | tee -a $O/summary.txt
grep "TOTAL" $O/stress_e4.log
timeout 600 python scripts/stress_s5.py 100 > $O/stress_s5.log 2>&1; echo "stress s5 rc=$?" | tee -a $O/summary.txt
grep "TOTAL" $O/stress_s5.log
for i in 1 2; do
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench$i.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench$i.json')); print('value', round(d['value']), 'us/step', round(1000*d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],4), round(d['roofline']['us_per_launch'],2))
k=d['kernels']
for n in ('stft_X','stft_X_mag_phase','istft'): print(' ', n, round(k[n]['us'],2), round(k[n]['frac'],3), '| b256', round(k['batch256'][n]['us'],1), round(k['batch256'][n]['frac'],3))
r=k['reference_default_geometry']
for n in ('explain','stft_X','stft_X_mag_phase','istft'): print('  refdef', n, round(r[n]['us'],2), round(r[n]['frac'],3))
print('  mel', k['mel_frontend']['us'], 'b1024', json.dumps(k.get('batch1024'))[:300])
"
done
