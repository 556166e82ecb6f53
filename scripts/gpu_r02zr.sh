#!/bin/bash
# round-2 GPU session ZR: fused mel front-end on the generation-3 transform, balanced slot ranges
cd "$(dirname "$0")/.."
O=gpurun_out/r02zr; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -x -q -m gpu -k "mel" > $O/pytest_mel.log 2>&1; echo "pytest mel rc=$?" | tee -a $O/summary.txt
tail -15 $O/pytest_mel.log
timeout 300 python scripts/mel_bench.py --tag hifigan_b64 > $O/mel.jsonl 2> $O/mel.err; echo "mel bench rc=$?" | tee -a $O/summary.txt
timeout 300 python scripts/mel_bench.py --default --n 80000 --tag default_b64 >> $O/mel.jsonl 2>> $O/mel.err
timeout 300 python scripts/mel_bench.py --batch 256 --pool 4 --tag hifigan_b256 >> $O/mel.jsonl 2>> $O/mel.err
cat $O/mel.jsonl; tail -5 $O/mel.err
