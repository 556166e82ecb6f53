#!/bin/bash
# round-2 GPU session ZD: ncu --set full of the HBM-bound 128-channel residual conv (conv1d_slab2, 4.3 GB per launch at 256 clips)
cd "$(dirname "$0")/.."
O=gpurun_out/r02zd; mkdir -p $O
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"conv1d_slab2" -s 51 -c 2 -f -o $O/prof_slab2 python scripts/bench_vocoder.py --batch 256 --iters 1 > $O/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
ncu -i $O/prof_slab2.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02zd/raw.csv')) if r]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__m_l1tex2xbar_write_bytes.sum','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum','sm__inst_executed_pipe_uniform.sum','sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed']
keys+= [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
for r in rows[2:]:
    print(r[ix['Kernel Name']][:60], r[ix['Grid Size']])
    for k in keys:
        if k in ix: print('   ',k.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio',''), rows[1][ix[k]], r[ix[k]])
PY
