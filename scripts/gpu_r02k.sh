#!/bin/bash
# round-2 GPU session K: why is the streaming iSTFT slower than generation 3?  ncu of both + one-CTA-per-SM A/B
cd "$(dirname "$0")/.."
O=gpurun_out/r02k; mkdir -p $O
cat > /tmp/prof_i4.py <<'PY'
import importlib, sys, torch
sys.path.insert(0, ".")
pkg = importlib.import_module("xai-audio-deepfakes_b200"); ops = pkg.ops
g = torch.Generator(device="cuda").manual_seed(0)
wav = 0.1 * torch.randn(64, 64000, generator=g, device="cuda")
X, _, _ = ops.stft(wav, 512, 160, 512, want_mag=False, want_phase=False)
for _ in range(3):
    y = ops.istft(X, 512, 160, 512, length=64000)
torch.cuda.synchronize(); print("ok")
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"istft4" -s 2 -c 1 -f -o $O/prof_i4 python /tmp/prof_i4.py > $O/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
cp xai-audio-deepfakes_b200/libaddvisor_sm100.so $O/lib_i4.so
ADV_NVCC_EXTRA=-DADV_ISTFT4_CTAS=1 python -c "
import importlib; pkg = importlib.import_module('xai-audio-deepfakes_b200'); pkg._lib.build(force=True)" > $O/rebuild.log 2>&1; echo "rebuild rc=$?" | tee -a $O/summary.txt
timeout 300 python scripts/kbench.py istft --tag i4_1cta_b64 > $O/kbench.jsonl 2> $O/kbench.err
timeout 300 python scripts/kbench.py istft --batch 256 --pool 4 --tag i4_1cta_b256 >> $O/kbench.jsonl 2>> $O/kbench.err
cut -c1-300 $O/kbench.jsonl
