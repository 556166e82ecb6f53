#!/bin/bash
# round-2 GPU session ZN: ncu --set full of the n_fft 1024 streaming kernels (explain5 / istft5) with source-level counts
cd "$(dirname "$0")/.."
O=gpurun_out/r02zn; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"explain5_kernel|istft5_kernel" -s 4 -c 2 -f -o $O/prof_s5 python scripts/prof_explain1024.py > $O/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
ncu -i $O/prof_s5.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $O/prof_s5.ncu-rep --page source --csv -k regex:explain5 > $O/src_e5.csv 2>/dev/null
cp xai-audio-deepfakes_b200/libaddvisor_sm100.so $O/lib.so
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02zn/raw.csv')) if r]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','sm__cycles_active.avg','sm__cycles_elapsed.max','launch__registers_per_thread','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']
keys+= [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
for r in rows[2:]:
    print(r[ix['Kernel Name']][:60])
    for k in keys:
        if k in ix: print('   ',k.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio',''), rows[1][ix[k]], r[ix[k]])
PY
