#!/bin/bash
# round-2 GPU session ZB: per-launch list of the HiFi-GAN generator at 256 clips (reflect padding): duration, tensor-pipe
# activity, DRAM throughput, L2 bytes
cd "$(dirname "$0")/.."
O=gpurun_out/r02zb; mkdir -p $O
timeout 600 python scripts/bench_vocoder.py --batch 256 --iters 5 > $O/voc.json 2> $O/voc.err; cat $O/voc.json | cut -c1-400
timeout 1500 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 160 -c 160 --csv --log-file $O/vocoder_launches_b256.csv python scripts/bench_vocoder.py --batch 256 --iters 1 > $O/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
tail -2 $O/ncu.log
