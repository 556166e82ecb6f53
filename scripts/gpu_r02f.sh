#!/bin/bash
# round-2 GPU session F: explain4 with 64-bit exchanges + single mask tile; register cap A/B (112 vs 128); ncu
cd "$(dirname "$0")/.."
O=gpurun_out/r02f; mkdir -p $O
timeout 600 python scripts/stress_e4.py 600 > $O/stress.log 2>&1; echo "stress rc=$?" | tee -a $O/summary.txt
grep -v "^MISMATCH\|per-row\|stats diff" $O/stress.log | tail -7
timeout 900 python -m pytest tests/test_gpu_explain4.py tests/test_gpu_parity.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest.log
K="timeout 300 python scripts/kbench.py"
$K explain --tag r112_b64 > $O/kbench.jsonl 2> $O/kbench.err
$K explain --batch 256 --pool 4 --tag r112_b256 >> $O/kbench.jsonl 2>> $O/kbench.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_r112.json 2> $O/bench.err; echo "bench112 rc=$?" | tee -a $O/summary.txt
cat > /tmp/prof_e4.py <<'PY'
import importlib, sys, torch
sys.path.insert(0, ".")
pkg = importlib.import_module("xai-audio-deepfakes_b200"); ops = pkg.ops
g = torch.Generator(device="cuda").manual_seed(0)
wav = 0.1 * torch.randn(64, 64000, generator=g, device="cuda"); mask = torch.rand(64, 257, 401, generator=g, device="cuda")
for _ in range(3):
    rel, irr = ops.explain(wav, mask, 512, 160, 512, length=64000, normalize=True)
torch.cuda.synchronize(); print("ok")
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"explain4" -s 2 -c 1 -f -o $O/prof_e4 python /tmp/prof_e4.py > $O/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
cp xai-audio-deepfakes_b200/libaddvisor_sm100.so $O/lib_r112.so
ADV_NVCC_EXTRA=-DADV_EXPLAIN4_MAXREG=128 python -c "
import importlib; pkg = importlib.import_module('xai-audio-deepfakes_b200'); pkg._lib.build(force=True)" > $O/rebuild.log 2>&1; echo "rebuild rc=$?" | tee -a $O/summary.txt
$K explain --tag r128_b64 >> $O/kbench.jsonl 2>> $O/kbench.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_r128.json 2>> $O/bench.err; echo "bench128 rc=$?" | tee -a $O/summary.txt
cut -c1-330 $O/kbench.jsonl
for f in $O/bench_r112.json $O/bench_r128.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', 'value', round(d['value']), 'ms/step', d['ms_per_step'], 'burst', d['run']['burst_us_per_step'], 'explain us', d['roofline']['us_per_launch'], 'e2e', round(d['e2e']['value']))"; done
