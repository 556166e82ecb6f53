#!/bin/bash
# round-2 GPU session ZX: same-box A/B of explain4_kernel's mask refill group (16 / 8 / 4 / 2 warps) and of the position of the
# tail-empty wait (variant libraries built by scripts/build_variant.sh, loaded through ADV_LIB_PATH)
cd "$(dirname "$0")/.."
O=gpurun_out/${OUT:-r02zx}; mkdir -p $O
P=$PWD/xai-audio-deepfakes_b200
for v in ${VARIANTS:-default}; do
  if [ $v = default ]; then unset ADV_LIB_PATH; else export ADV_LIB_PATH=$P/libaddvisor_sm100.$v.so; fi
  timeout 600 python -m pytest tests/test_gpu_explain4.py tests/test_gpu_parity.py tests/test_gpu_end_to_end.py tests/test_gpu_stream1024.py tests/test_gpu_tensorcore.py -x -q -m gpu > $O/pytest_$v.log 2>&1; rc=$?
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.load(open('$O/bench_$v.json')); print(json.dumps({'variant':'$v','pytest_rc':$rc,'value':round(d['value']),'us_per_step':round(1000*d['ms_per_step'],2),'explain_us':round(d['roofline']['us_per_launch'],2),'burst':d['run']['burst_us_per_step'],'stft_X':round(d['kernels']['stft_X']['us'],2),'stft_Xmp':round(d['kernels']['stft_X_mag_phase']['us'],2),'istft':round(d['kernels']['istft']['us'],2),'refdef':{k:round(v['us'],1) for k,v in d['kernels']['reference_default_geometry'].items() if isinstance(v,dict)}}))" | tee -a $O/ab.jsonl
done
