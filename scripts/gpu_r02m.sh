#!/bin/bash
# round-2 GPU session M: streaming iSTFT with register prefetch; STFT at four CTAs per SM
cd "$(dirname "$0")/.."
O=gpurun_out/r02m; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_explain4.py tests/test_gpu_parity.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest.log
K="timeout 300 python scripts/kbench.py"
{
$K istft stft stft3 --tag i4pre_b64
$K istft stft stft3 --batch 256 --pool 4 --tag i4pre_b256
$K istft --hop 128 --tag i4pre_hop128
$K istft --hop 256 --tag i4pre_hop256
} > $O/kbench.jsonl 2> $O/kbench.err
timeout 600 python scripts/stress_e4.py 300 > $O/stress.log 2>&1; grep "TOTAL" $O/stress.log
ADV_NVCC_EXTRA="-DADV_STFT3_VEC_CTAS=4" python -c "
import importlib; pkg = importlib.import_module('xai-audio-deepfakes_b200'); pkg._lib.build(force=True)" > $O/rebuild.log 2>&1; echo "rebuild rc=$?" | tee -a $O/summary.txt
$K stft stft3 --tag stft4ctas_b64 >> $O/kbench.jsonl 2>> $O/kbench.err
$K stft stft3 --batch 256 --pool 4 --tag stft4ctas_b256 >> $O/kbench.jsonl 2>> $O/kbench.err
cut -c1-700 $O/kbench.jsonl
