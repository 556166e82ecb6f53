#!/bin/bash
# round-2 GPU session ZZD (8 GPUs): driver-style bench at N = 8, 4, 2, 1 with the final kernels
cd "$(dirname "$0")/.."
O=gpurun_out/r02zzd; mkdir -p $O
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${n}gpu.json 2> $O/bench$n.err; echo "bench$n rc=$?" | tee -a $O/summary.txt
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench1.err; echo "bench1 rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
for f in ('bench_1gpu','bench_2gpu','bench_4gpu','bench_8gpu'):
    try:
        d=json.load(open('gpurun_out/r02zzd/%s.json'%f)); print(f, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'n', d['n_gpus'], 'us/step', round(1000*d.get('ms_per_step'),2), d['run'].get('host_affinity'))
    except Exception as e: print(f, 'ERR', e)
PY
