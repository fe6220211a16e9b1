#!/bin/bash
# atomic in-tree build of the two libraries (a GPU-box snapshot taken meanwhile never sees a half-written .so)
set -e
R="$(cd "$(dirname "$0")/.." && pwd)"
cd "$R/apus_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -I../../include -I. -shared -o /tmp/libapus_gpu.so.new apus_kernels.cu apus_engine.cu -lrt
mv /tmp/libapus_gpu.so.new "$R/apus_b200/libapus_gpu.so"
gcc -O2 -g -std=gnu99 -fPIC -Wall -shared -I"$R/include" -o /tmp/libapus_dare.so.new dare_entry.c -L"$R/apus_b200" -lapus_gpu -Wl,-rpath,'$ORIGIN' -lpthread
mv /tmp/libapus_dare.so.new "$R/apus_b200/libapus_dare.so"
echo built
