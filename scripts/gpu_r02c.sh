#!/bin/bash
# round-2 GPU session C (re-run after the container was re-created): parity, gen3 vs gen2 kernel timings, configs[0],
# bench (both arms), launch list + --set full capture of the gen3 kernels
cd "$(dirname "$0")/.."
O=gpurun_out/r02c; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gen3.log 2>&1; echo "pytest gen3 rc=$?" | tee -a $O/summary.txt
K="timeout 300 python scripts/kbench.py"
{
$K --tag gen3_cfg2_b64
ADV_STFT3_VEC=0 $K stft stft3 --tag gen3planar_cfg2_b64
ADV_GEN3=0 $K --tag gen2_cfg2_b64
$K --batch 256 --pool 4 --tag gen3_cfg2_b256
ADV_GEN3=0 $K --batch 256 --pool 4 --tag gen2_cfg2_b256
$K --nfft 1024 --hop 322 --n 80000 --pool 8 --tag gen3_refdef_b64
ADV_GEN3=0 $K --nfft 1024 --hop 322 --n 80000 --pool 8 --tag gen2_refdef_b64
$K --nfft 1024 --hop 256 --n 64000 --win hann --winlen 1024 --pool 8 --tag gen3_hifigan_geom
} > $O/kbench.jsonl 2> $O/kbench.err
timeout 600 python scripts/cfg1_wavs.py > $O/cfg1.json 2> $O/cfg1.err; echo "cfg1 rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?" | tee -a $O/summary.txt
timeout 600 python scripts/prof_kernels3.py > $O/prof_plain.log 2>&1; echo "prof plain rc=$?" | tee -a $O/summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"explain3|istft3|stft3" -s 14 -c 7 -f -o $O/prof python scripts/prof_kernels3.py > $O/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 5 > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
ls -la $O
tail -3 $O/pytest_gen3.log; cut -c1-700 $O/kbench.jsonl; cut -c1-3000 $O/bench.json
