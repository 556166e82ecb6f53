#!/bin/bash
# round-2 GPU session ZO (8 GPUs): driver-style bench at N = 8 and 4, BASELINE configs[3] (100 000 clips batch-sharded) at N = 8
cd "$(dirname "$0")/.."
O=gpurun_out/r02zo; mkdir -p $O
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > $O/bench_${n}gpu.json 2> $O/bench$n.err; echo "bench$n rc=$?" | tee -a $O/summary.txt
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 scripts/cfg4_eval.py --clips 100000 > $O/cfg4_8gpu.json 2> $O/cfg4.err; echo "cfg4 rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
for f in ('bench_8gpu','bench_4gpu'):
    try:
        d=json.load(open('gpurun_out/r02zo/%s.json'%f)); print(f, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'n', d['n_gpus'], 'ms/step', d.get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
print(open('gpurun_out/r02zo/cfg4_8gpu.json').read()[:600])
PY
