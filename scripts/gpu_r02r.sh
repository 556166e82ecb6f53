#!/bin/bash
# round-2 GPU session R: batched load_audio (decode + resample + pad) parity; whole GPU suite
cd "$(dirname "$0")/.."
O=gpurun_out/r02r; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_load_audio.py -x -q -m gpu > $O/pytest_load.log 2>&1; echo "pytest load rc=$?" | tee -a $O/summary.txt
tail -15 $O/pytest_load.log
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_all.log
