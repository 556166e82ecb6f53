#!/bin/bash
# round-2 GPU session ZK: residual TMA as a template specialisation
cd "$(dirname "$0")/.."
O=gpurun_out/r02zk; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -x -q -m gpu > $O/pytest_tc.log 2>&1; echo "pytest tc rc=$?" | tee -a $O/summary.txt
tail -4 $O/pytest_tc.log
for e in tma tma_st direct; do
  timeout 600 python scripts/bench_vocoder.py --batch 256 --iters 5 --epilogue $e > $O/voc_$e.json 2>> $O/voc.err; cut -c1-330 $O/voc_$e.json
done
tail -3 $O/voc.err
timeout 1500 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 160 -c 160 --csv --log-file $O/vocoder_launches_b256.csv python scripts/bench_vocoder.py --batch 256 --iters 1 > $O/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
