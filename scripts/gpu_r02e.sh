#!/bin/bash
# round-2 GPU session E: after the slice write-after-read fix - stress, parity, timings
cd "$(dirname "$0")/.."
O=gpurun_out/r02e; mkdir -p $O
timeout 900 python scripts/stress_e4.py 1500 > $O/stress.log 2>&1; echo "stress rc=$?" | tee -a $O/summary.txt
grep -v "^MISMATCH\|per-row\|stats diff" $O/stress.log | tail -8
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest.log
K="timeout 300 python scripts/kbench.py"
{
$K --tag gen4_cfg2_b64
$K --batch 256 --pool 4 --tag gen4_cfg2_b256
$K --nfft 1024 --hop 322 --n 80000 --pool 8 --tag gen3_refdef_b64
} > $O/kbench.jsonl 2> $O/kbench.err
cut -c1-700 $O/kbench.jsonl
