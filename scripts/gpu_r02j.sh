#!/bin/bash
# round-2 GPU session J: streaming iSTFT (istft4) parity / stress / timing; STFT residency A/B
cd "$(dirname "$0")/.."
O=gpurun_out/r02j; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -8 $O/pytest.log
timeout 900 python scripts/stress_e4.py 500 > $O/stress.log 2>&1; echo "stress rc=$?" | tee -a $O/summary.txt
grep -v "^MISMATCH\|per-row\|stats diff" $O/stress.log | tail -11
K="timeout 300 python scripts/kbench.py"
{
$K istft --tag i4_cfg2_b64
ADV_GEN4=0 $K istft --tag i3_cfg2_b64
$K istft --batch 256 --pool 4 --tag i4_cfg2_b256
$K istft --hop 128 --tag i4_hop128
$K istft --hop 256 --tag i4_hop256
} > $O/kbench.jsonl 2> $O/kbench.err
cut -c1-300 $O/kbench.jsonl
