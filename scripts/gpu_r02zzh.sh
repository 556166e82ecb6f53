#!/bin/bash
# round-2 GPU session ZZH: soak of the final kernels - race hunts at higher iteration counts, the GPU suite twice
cd "$(dirname "$0")/.."
O=gpurun_out/r02zzh; mkdir -p $O
timeout 900 python scripts/stress_e4.py 500 > $O/stress_e4.log 2>&1; echo "stress e4 rc=$?" | tee -a $O/summary.txt
grep -v "^MISMATCH\|per-row\|stats diff" $O/stress_e4.log | tail -8
timeout 900 python scripts/stress_s5.py 250 > $O/stress_s5.log 2>&1; echo "stress s5 rc=$?" | tee -a $O/summary.txt
tail -4 $O/stress_s5.log
for i in 1 2; do
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $O/pytest_all$i.log 2>&1; echo "pytest all $i rc=$?" | tee -a $O/summary.txt
tail -2 $O/pytest_all$i.log | head -1
done
