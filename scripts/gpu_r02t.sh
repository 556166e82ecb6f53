#!/bin/bash
# round-2 GPU session T: dynamic item scheduling in the STFT + programmatic dependent launch on by default
cd "$(dirname "$0")/.."
O=gpurun_out/r02t; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_all.log
timeout 600 python scripts/stress_e4.py 300 > $O/stress.log 2>&1; echo "stress rc=$?" | tee -a $O/summary.txt
grep -v "^MISMATCH\|per-row\|stats diff" $O/stress.log | tail -4
K="timeout 300 python scripts/kbench.py"
$K --tag dyn_pdl_b64 > $O/kbench.jsonl 2> $O/kbench.err
$K --batch 256 --pool 4 --tag dyn_pdl_b256 >> $O/kbench.jsonl 2>> $O/kbench.err
$K stft stft3 istft --nfft 1024 --hop 322 --n 80000 --tag dyn_pdl_refdef >> $O/kbench.jsonl 2>> $O/kbench.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02t/kbench.jsonl'):
    d=json.loads(ln); print(d['tag'], {k:(round(v['us'],2), round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict)})
PY
tail -3 $O/kbench.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch'], d['roofline']['traffic']); k=d['kernels']; print({n:(round(k[n]['us'],2), round(k[n]['frac'],3)) for n in ('stft_X','stft_X_mag_phase','istft')}); print(k['batch256']); print(k['reference_default_geometry'])"
