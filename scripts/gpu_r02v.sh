#!/bin/bash
# round-2 GPU session V: STFT A/B - static round-robin vs grouped draw counters (padded), each with PDL off / on
cd "$(dirname "$0")/.."
O=gpurun_out/r02v; mkdir -p $O
K="timeout 300 python scripts/kbench.py stft stft3"
run() {
  for pdl in 0 1; do
    $K --pdl $pdl --tag $1_pdl${pdl}_b64 >> $O/kbench.jsonl 2>> $O/kbench.err
    $K --pdl $pdl --batch 256 --pool 4 --tag $1_pdl${pdl}_b256 >> $O/kbench.jsonl 2>> $O/kbench.err
    $K --pdl $pdl --batch 1024 --pool 2 --reps 10 --tag $1_pdl${pdl}_b1024 >> $O/kbench.jsonl 2>> $O/kbench.err
  done
}
run static
ADV_NVCC_EXTRA=-DADV_STFT3_DYN=1 python -c "
import importlib; pkg = importlib.import_module('xai-audio-deepfakes_b200'); pkg._lib.build(force=True)" > $O/rebuild.log 2>&1; echo "rebuild rc=$?" | tee -a $O/summary.txt
run dyn
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k stft > $O/pytest_dyn.log 2>&1; echo "pytest dyn rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
for ln in open('gpurun_out/r02v/kbench.jsonl'):
    d=json.loads(ln); print(d['tag'], {k:(round(v['us'],2), round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict)})
PY
tail -3 $O/kbench.err
