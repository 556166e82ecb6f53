#!/bin/bash
# round-2 GPU session Q (8 GPUs): BASELINE configs[4] as specified - real integrated gradients (50 Gauss-Legendre steps through
# Wav2vec2LogReg on a seeded classifier) on 30 s clips feeding td_mask -> 3 x compute_stft -> LMAC sums, one all-reduce
cd "$(dirname "$0")/.."
O=gpurun_out/r02q; mkdir -p $O
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  scripts/cfg5_longform.py --ig 2 > $O/cfg5_ig_${N}gpu.json 2> $O/cfg5.err; echo "cfg5 x$N rc=$?" | tee -a $O/summary.txt
cut -c1-1800 $O/cfg5_ig_${N}gpu.json; tail -3 $O/cfg5.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_${N}gpu.json 2> $O/bench.err; echo "bench x$N rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench_${N}gpu.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['n_gpus'])"
