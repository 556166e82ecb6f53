#!/bin/bash
# round-2 GPU session A: parity of the generation-3 kernels, A/B timings against generation 2, first bench line
cd "$(dirname "$0")/.."
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gen3.log 2>&1; echo "pytest gen3 rc=$?" | tee -a $O/summary.txt
ADV_GEN3=0 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $O/pytest_gen2.log 2>&1; echo "pytest gen2 rc=$?" | tee -a $O/summary.txt
K="timeout 300 python scripts/kbench.py"
{
$K --tag gen3_cfg2_b64
ADV_STFT3_VEC=0 $K stft stft3 --tag gen3planar_cfg2_b64
ADV_GEN3=0 $K --tag gen2_cfg2_b64
$K --batch 256 --pool 4 --tag gen3_cfg2_b256
ADV_STFT3_VEC=0 $K stft stft3 --batch 256 --pool 4 --tag gen3planar_cfg2_b256
ADV_GEN3=0 $K --batch 256 --pool 4 --tag gen2_cfg2_b256
$K --nfft 1024 --hop 322 --n 80000 --pool 8 --tag gen3_refdef_b64
ADV_STFT3_VEC=0 $K stft stft3 --nfft 1024 --hop 322 --n 80000 --pool 8 --tag gen3planar_refdef_b64
ADV_GEN3=0 $K --nfft 1024 --hop 322 --n 80000 --pool 8 --tag gen2_refdef_b64
$K --nfft 1024 --hop 256 --n 64000 --win hann --winlen 1024 --pool 8 --tag gen3_hifigan_geom
ADV_GEN3=0 $K --nfft 1024 --hop 256 --n 64000 --win hann --winlen 1024 --pool 8 --tag gen2_hifigan_geom
} > $O/kbench.jsonl 2> $O/kbench.err
timeout 600 python scripts/cfg1_wavs.py > $O/cfg1.json 2> $O/cfg1.err; echo "cfg1 rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py --steps 64 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gen3.log; tail -3 $O/pytest_gen2.log; cat $O/kbench.jsonl | cut -c1-600
