#!/bin/bash
# round-2 GPU session D: generation-4 streaming explain kernel - parity, racecheck on a small case, timings, ncu
cd "$(dirname "$0")/.."
O=gpurun_out/r02d; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_explain4.py -x -q > $O/pytest_e4.log 2>&1; echo "pytest e4 rc=$?" | tee -a $O/summary.txt
tail -15 $O/pytest_e4.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_end_to_end.py tests/test_gpu_sharded.py -x -q > $O/pytest_rest.log 2>&1; echo "pytest rest rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_rest.log
cat > /tmp/small_e4.py <<'PY'
import importlib, sys, torch
sys.path.insert(0, ".")
pkg = importlib.import_module("xai-audio-deepfakes_b200"); ops = pkg.ops
g = torch.Generator().manual_seed(0)
B, n, hop = 20, 6400, 160
wav = 0.1 * torch.randn(B, n, generator=g); mask = torch.rand(B, 257, 1 + n // hop, generator=g)
rel, irr = ops.explain(wav, mask, 512, hop, 512, length=n, normalize=True)
torch.cuda.synchronize(); print("ok", float(rel.abs().max()))
PY
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python /tmp/small_e4.py > $O/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/summary.txt
tail -8 $O/racecheck.log
timeout 600 compute-sanitizer --tool memcheck python /tmp/small_e4.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/summary.txt
tail -4 $O/memcheck.log
K="timeout 300 python scripts/kbench.py"
{
$K explain --tag gen4_cfg2_b64
ADV_GEN4=0 $K explain --tag gen3_cfg2_b64
$K explain --batch 256 --pool 4 --tag gen4_cfg2_b256
$K explain --hop 128 --tag gen4_hop128
$K explain --hop 256 --tag gen4_hop256
} > $O/kbench.jsonl 2> $O/kbench.err
cut -c1-400 $O/kbench.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
cut -c1-1200 $O/bench.json
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
