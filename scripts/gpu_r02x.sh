#!/bin/bash
# round-2 GPU session X: streaming fused explain for n_fft 1024 (explain5) - parity + timing on the reference-default geometry
cd "$(dirname "$0")/.."
O=gpurun_out/r02x; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_stream1024.py -x -q -m gpu > $O/pytest_s5.log 2>&1; echo "pytest s5 rc=$?" | tee -a $O/summary.txt
tail -12 $O/pytest_s5.log
K="timeout 300 python scripts/kbench.py"
$K explain istft --nfft 1024 --hop 322 --n 80000 --tag e5_refdef_b64 > $O/kbench.jsonl 2> $O/kbench.err
$K explain --nfft 1024 --hop 322 --n 80000 --batch 256 --pool 4 --tag e5_refdef_b256 >> $O/kbench.jsonl 2>> $O/kbench.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02x/kbench.jsonl'):
    d=json.loads(ln); print(d['tag'], {k:(round(v['us'],2), round(v['frac'],3), v.get('err')) for k,v in d.items() if isinstance(v,dict)})
PY
tail -3 $O/kbench.err
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_all.log
