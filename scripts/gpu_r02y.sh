#!/bin/bash
# round-2 GPU session Y: does de-correlating the output arrays' DRAM addresses change the 3-output STFT? + bench line
cd "$(dirname "$0")/.."
O=gpurun_out/r02y; mkdir -p $O
K="timeout 300 python scripts/kbench.py stft3"
for skew in 0 1088 33000 262144; do
  ADV_STFT_SKEW=$skew $K --tag skew${skew}_b64 >> $O/kbench.jsonl 2>> $O/kbench.err
  ADV_STFT_SKEW=$skew $K --batch 256 --pool 4 --tag skew${skew}_b256 >> $O/kbench.jsonl 2>> $O/kbench.err
  ADV_STFT_SKEW=$skew $K --nfft 1024 --hop 322 --n 80000 --tag skew${skew}_refdef >> $O/kbench.jsonl 2>> $O/kbench.err
done
python - <<'PY'
import json
for ln in open('gpurun_out/r02y/kbench.jsonl'):
    d=json.loads(ln); print(d['tag'], {k:(round(v['us'],2), round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict)})
PY
tail -3 $O/kbench.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch'], d['roofline']['traffic']); k=d['kernels']; print({n:(round(k[n]['us'],2), round(k[n]['frac'],3)) for n in ('stft_X','stft_X_mag_phase','istft')}); print(k['batch256']); print(k['reference_default_geometry']); print(k['mel_frontend'])"
