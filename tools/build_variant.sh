#!/bin/bash
# tools/build_variant.sh NAME "-DFLAGS": builds gpurun_ab/libswcgpu_NAME.so with extra flags for inflate_lut.cu / inflate.cu (A/B runs via SWCGPU_SO)
set -e
cd /root/repo/swcompression_b200/csrc
name=$1; flags=$2
mkdir -p /tmp/w/var_$name
objs=""
for f in *.cu; do
  o=build/${f%.cu}.o
  if [ "$f" = inflate_lut.cu ] || [ "$f" = inflate.cu ]; then
    o=/tmp/w/var_$name/${f%.cu}.o
    /usr/local/cuda/bin/nvcc $flags -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v -c $f -o $o 2> /tmp/w/var_$name/${f%.cu}.log
    grep -A2 "inflate_lut_kernel" /tmp/w/var_$name/${f%.cu}.log | grep -E "registers|spill" || true
  fi
  objs="$objs $o"
done
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /root/repo/gpurun_ab/libswcgpu_$name.so $objs -lcudart
echo built $name
