#!/bin/bash
# round-2 GPU session Z: one-frame-per-warp STFT for n_fft 1024 (stft5) - parity + timing
cd "$(dirname "$0")/.."
O=gpurun_out/r02z; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -4 $O/pytest_all.log
K="timeout 300 python scripts/kbench.py"
$K stft stft3 --nfft 1024 --hop 322 --n 80000 --tag f5_refdef_b64 > $O/kbench.jsonl 2> $O/kbench.err
$K stft stft3 --nfft 1024 --hop 322 --n 80000 --batch 256 --pool 4 --tag f5_refdef_b256 >> $O/kbench.jsonl 2>> $O/kbench.err
$K stft stft3 --nfft 1024 --hop 256 --win hann --winlen 1024 --tag f5_hifigan_b64 >> $O/kbench.jsonl 2>> $O/kbench.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02z/kbench.jsonl'):
    d=json.loads(ln); print(d['tag'], {k:(round(v['us'],2), round(v['frac'],3), v.get('err', v.get('err_mag'))) for k,v in d.items() if isinstance(v,dict)})
PY
tail -3 $O/kbench.err
timeout 300 python scripts/cfg5_longform.py > $O/cfg5.json 2> $O/cfg5.err; cut -c1-700 $O/cfg5.json
