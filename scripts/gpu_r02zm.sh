#!/bin/bash
# round-2 GPU session ZM (2 GPUs): smoke, bench at N = 1 and N = 2 with the driver's flags, reference arm
cd "$(dirname "$0")/.."
O=gpurun_out/r02zm; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt; tail -2 $O/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench1.err; echo "bench1 rc=$?" | tee -a $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench2.err; echo "bench2 rc=$?" | tee -a $O/summary.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/benchr.err; echo "benchref rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
for f in ('bench_1gpu','bench_2gpu','bench_reference'):
    try:
        d=json.load(open('gpurun_out/r02zm/%s.json'%f))
        print(f, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'n', d['n_gpus'], 'ms/step', d.get('ms_per_step'), 'clocks', d.get('clocks',{}).get('sm_mhz'), d.get('clocks',{}).get('reasons'))
        if 'kernels' in d:
            k=d['kernels']; print('  ', {n:(round(k[n]['us'],2), round(k[n]['frac'],3)) for n in ('stft_X','stft_X_mag_phase','istft')}, 'roofline', round(d['roofline']['frac'],4), round(d['roofline']['us_per_launch'],2)); print('   b256', {n:(round(v['us'],1), round(v['frac'],3)) for n,v in k['batch256'].items() if isinstance(v,dict)}); print('   refdef', {n:(round(v['us'],1), round(v['frac'],3)) for n,v in k['reference_default_geometry'].items() if isinstance(v,dict)}); print('   voc', d['vocoder'].get('clips_per_s'), d['vocoder'].get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
