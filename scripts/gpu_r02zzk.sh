#!/bin/bash
# round-2 GPU session ZZK: per-instruction profile of the reference-default STFT with magnitude and phase (stft_w_kernel<1024>)
cd "$(dirname "$0")/.."
O=gpurun_out/r02zzk; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stft_w_kernel|stft5_kernel" -c 4 -f -o $O/prof python scripts/prof_stft1024.py > $O/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
ncu -i $O/prof.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $O/prof.ncu-rep --page source --csv > $O/source.csv 2>/dev/null
rm -f $O/prof.ncu-rep
ls -la $O
