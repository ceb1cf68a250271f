#!/bin/bash
# A/B builds of libprb_b200.so with compile-time knobs: tools/build_variants.sh NAME "-DFOO=1 -DBAR=2" ...
# -> roboticsplayroompybullet_b200/variants/libprb_b200_NAME.so ; select with PRB_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
mkdir -p roboticsplayroompybullet_b200/variants
while [ $# -ge 2 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $2 \
       -o roboticsplayroompybullet_b200/variants/libprb_b200_$1.so roboticsplayroompybullet_b200/csrc/prb_capi.cu &
  shift 2
done
wait
ls -la roboticsplayroompybullet_b200/variants/
