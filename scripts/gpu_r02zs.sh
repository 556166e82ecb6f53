#!/bin/bash
# round-2 GPU session ZS: race hunt on the n_fft 1024 streaming kernels and the fused mel front-end
cd "$(dirname "$0")/.."
O=gpurun_out/r02zs; mkdir -p $O
timeout 900 python scripts/stress_s5.py 300 > $O/stress_s5.log 2>&1; echo "stress s5 rc=$?" | tee -a $O/summary.txt
tail -8 $O/stress_s5.log
timeout 600 python scripts/stress_e4.py 200 > $O/stress_e4.log 2>&1; echo "stress e4 rc=$?" | tee -a $O/summary.txt
grep -v "^MISMATCH\|per-row\|stats diff" $O/stress_e4.log | tail -3
