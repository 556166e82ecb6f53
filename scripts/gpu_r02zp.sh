#!/bin/bash
# round-2 GPU session ZP: stft5 parity cases; whole suite
cd "$(dirname "$0")/.."
O=gpurun_out/r02zp; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_stream1024.py -x -q -m gpu -k stft1024 > $O/pytest_f5.log 2>&1; echo "pytest f5 rc=$?" | tee -a $O/summary.txt
tail -12 $O/pytest_f5.log
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_all.log
