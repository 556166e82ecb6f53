#!/bin/bash
# round-2 GPU session ZZU: stft3_kernel loading the 16 + HS distinct sample rows of a frame pair once - parity + kbench A/B
cd "$(dirname "$0")/.."
O=gpurun_out/r02zzu; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_end_to_end.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -2 $O/pytest.log | head -1
for v in default noshare default noshare; do
  if [ $v = default ]; then unset ADV_LIB_PATH; else export ADV_LIB_PATH=$PWD/xai-audio-deepfakes_b200/libaddvisor_sm100.$v.so; fi
  timeout 200 python scripts/kbench.py stft stft3 --tag $v 2>/dev/null | tail -1 | cut -c1-420 | tee -a $O/kbench.jsonl
done
unset ADV_LIB_PATH
timeout 200 python scripts/kbench.py stft stft3 --hop 128 --tag default_hop128 2>/dev/null | tail -1 | cut -c1-300 | tee -a $O/kbench.jsonl
timeout 200 python scripts/kbench.py stft stft3 --hop 256 --win hann --tag default_hop256_hann 2>/dev/null | tail -1 | cut -c1-300 | tee -a $O/kbench.jsonl
