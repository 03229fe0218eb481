#!/bin/bash
# product library + the -DBMNAS_TIMELINE instrumented copy (scratch only)
set -e
cd /root/repo
python bm-nas_b200/build.py 2>&1 | tail -3
mkdir -p bm-nas_b200/build_tl
for f in bm-nas_b200/csrc/*.cu; do
  o=bm-nas_b200/build_tl/$(basename $f .cu).o
  if [ ! -f $o ] || [ $f -nt $o ] || [ bm-nas_b200/csrc/gemm_shared.cuh -nt $o ] || [ bm-nas_b200/csrc/common.cuh -nt $o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DBMNAS_TIMELINE -c $f -o $o &
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/ubench/libbmnas_tl.so bm-nas_b200/build_tl/*.o
ls -la bm-nas_b200/libbmnas_b200.so tools/ubench/libbmnas_tl.so
