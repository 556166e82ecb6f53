#!/bin/bash
# round-2 GPU session ZZM: per-instruction profile of the reference-default fused kernel and iSTFT (explain5 / istft5)
cd "$(dirname "$0")/.."
O=gpurun_out/r02zzm; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"explain5_kernel|istft5_kernel" -c 4 -f -o $O/prof python scripts/prof_explain1024.py > $O/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
ncu -i $O/prof.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $O/prof.ncu-rep --page source --csv > $O/source.csv 2>/dev/null
rm -f $O/prof.ncu-rep
tail -3 $O/ncu.log
