#!/bin/bash
# round-2 GPU session L: streaming iSTFT CTA shape (8 warps x 3 CTAs at 80 registers vs 8 x 2 at 120)
cd "$(dirname "$0")/.."
O=gpurun_out/r02l; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_explain4.py tests/test_gpu_parity.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest.log
K="timeout 300 python scripts/kbench.py"
{
$K istft --tag i4_8x3_b64
$K istft --batch 256 --pool 4 --tag i4_8x3_b256
$K istft --hop 128 --tag i4_8x3_hop128
$K istft --hop 256 --tag i4_8x3_hop256
} > $O/kbench.jsonl 2> $O/kbench.err
ADV_NVCC_EXTRA="-DADV_ISTFT4_CTAS=2" python -c "
import importlib; pkg = importlib.import_module('xai-audio-deepfakes_b200'); pkg._lib.build(force=True)" > $O/rebuild.log 2>&1; echo "rebuild rc=$?" | tee -a $O/summary.txt
$K istft --tag i4_8x2_b64 >> $O/kbench.jsonl 2>> $O/kbench.err
$K istft --batch 256 --pool 4 --tag i4_8x2_b256 >> $O/kbench.jsonl 2>> $O/kbench.err
cut -c1-300 $O/kbench.jsonl
