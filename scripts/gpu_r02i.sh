#!/bin/bash
# round-2 GPU session I: reflect-padded vocoder on the TMA kernels (halo layout)
cd "$(dirname "$0")/.."
O=gpurun_out/r02i; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_tensorcore.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -12 $O/pytest.log
timeout 600 python scripts/bench_vocoder.py > $O/voc.log 2>&1; tail -5 $O/voc.log
timeout 600 python scripts/bench_vocoder.py --batch 256 --iters 5 >> $O/voc.log 2>&1; tail -2 $O/voc.log | cut -c1-500
