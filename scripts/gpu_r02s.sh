#!/bin/bash
# round-2 GPU session S: bench line, launch list of the same command, --set full capture of the production kernels at
# 1 024 clips (write-back counted) -> profiles/traffic.json
cd "$(dirname "$0")/.."
O=gpurun_out/r02s; mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch']); print(json.dumps(d['kernels']['mel_frontend'])[:400])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_steps20.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"explain4_kernel|stft3_kernel|istft4_kernel|mel_fused_kernel" -f -o $O/prof_full python scripts/prof_traffic.py > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $O/summary.txt
tail -3 $O/ncu_full.log
ncu -i $O/prof_full.ncu-rep --page raw --csv > $O/raw_all.csv 2>/dev/null
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02s/raw_all.csv')) if r]
print(len(rows)-2, 'launches captured')
PY
