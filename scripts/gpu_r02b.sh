#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02b; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gen3.log 2>&1; echo "pytest gen3 rc=$?" | tee -a $O/summary.txt
ADV_GEN3=0 timeout 900 python -m pytest tests/test_gpu_parity.py -q > $O/pytest_gen2.log 2>&1; echo "pytest gen2 rc=$?" | tee -a $O/summary.txt
timeout 600 python scripts/prof_kernels3.py > $O/prof_plain.log 2>&1; echo "prof plain rc=$?" | tee -a $O/summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"explain3|istft3|stft3" -s 14 -c 7 -f -o $O/prof python scripts/prof_kernels3.py > $O/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
ls -la $O
tail -5 $O/pytest_gen3.log; tail -5 $O/pytest_gen2.log
