#!/bin/bash
# round-2 GPU session ZL: vocoder with the per-layer epilogue rule; full GPU suite; bench line
cd "$(dirname "$0")/.."
O=gpurun_out/r02zl; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_all.log
for e in tma direct; do
  timeout 600 python scripts/bench_vocoder.py --batch 256 --iters 5 --epilogue $e > $O/voc_$e.json 2>> $O/voc.err; cut -c1-300 $O/voc_$e.json
done
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch']); print(d['vocoder']); k=d['kernels']; print({n:(round(k[n]['us'],2), round(k[n]['frac'],3)) for n in ('stft_X','stft_X_mag_phase','istft')}); print(k['reference_default_geometry']); print(k['mel_frontend'])"
