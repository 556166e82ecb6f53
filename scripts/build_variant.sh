#!/bin/bash
# A/B build of the library: recompiles the named translation unit(s) with extra nvcc flags and links them with the objects
# of the default build (build/obj) into xai-audio-deepfakes_b200/libaddvisor_sm100.<name>.so; load it with ADV_LIB_PATH.
#   scripts/build_variant.sh gw4 "-DADV_E4_MASK_GW=4" transform4_kernels
set -e
cd "$(dirname "$0")/.."
name=$1; extra=$2; shift 2
P=xai-audio-deepfakes_b200
mkdir -p build/obj_$name
objs=""
for o in build/obj/*.o; do
  b=$(basename $o .o); use=$o
  for u in "$@"; do
    if [ "$b" = "$u" ]; then
      nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC $extra -c $P/csrc/$u.cu -o build/obj_$name/$u.o
      use=build/obj_$name/$u.o
    fi
  done
  objs="$objs $use"
done
nvcc -shared -o $P/libaddvisor_sm100.$name.so $objs
echo built $P/libaddvisor_sm100.$name.so
