#!/bin/bash
# round-2 GPU session U: STFT with grouped draw counters
cd "$(dirname "$0")/.."
O=gpurun_out/r02u; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_end_to_end.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -2 $O/pytest.log
K="timeout 300 python scripts/kbench.py"
$K stft stft3 --tag grp_b64 > $O/kbench.jsonl 2> $O/kbench.err
$K stft stft3 --batch 256 --pool 4 --tag grp_b256 >> $O/kbench.jsonl 2>> $O/kbench.err
$K stft stft3 --batch 1024 --pool 1 --tag grp_b1024 >> $O/kbench.jsonl 2>> $O/kbench.err
$K stft stft3 --batch 7 --n 16000 --tag grp_b7 >> $O/kbench.jsonl 2>> $O/kbench.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02u/kbench.jsonl'):
    d=json.loads(ln); print(d['tag'], {k:(round(v['us'],2), round(v['frac'],3), v.get('err', v.get('err_mag'))) for k,v in d.items() if isinstance(v,dict)})
PY
tail -3 $O/kbench.err
