#!/bin/bash
# Per-role wait-cycle totals of the GEMM kernels' first CTA for one forward pass of 1024 images.  Needs a trace build of the
# library (the product build carries no instrumentation):
#   cd vision-transformer-opencl_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DVIT_GEMM_TRACE \
#       -c csrc/engine.cu -o /tmp/engine_trace.o && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/libvit_b200_trace.so \
#       /tmp/engine_trace.o build/network_io.o build/results.o build/synth.o build/vit_cuda_adaptor.o -lm
L=vision-transformer-opencl_b200/lib
cp $L/libvit_b200.so /tmp/product.so
cp $L/libvit_b200_trace.so $L/libvit_b200.so
timeout 300 python tools/profile_one.py 1024 1 > gpurun_out/gemm_trace.log 2>&1
cp /tmp/product.so $L/libvit_b200.so
grep -c gemm_trace gpurun_out/gemm_trace.log
