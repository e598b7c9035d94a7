#!/bin/bash
# Per-role wait-cycle totals of the GEMM kernels' first CTA for one forward pass of 1024 images.  Needs the trace build of the
# library (`make -C vision-transformer-opencl_b200 trace`; the product build carries no instrumentation); the product library
# is put back afterwards.  Output: gpurun_out/gemm_trace.log (summarised in profiles/r2_gemm_trace.txt).
L=vision-transformer-opencl_b200/lib
cp $L/libvit_b200.so /tmp/product.so
cp $L/libvit_b200_trace.so $L/libvit_b200.so
timeout 300 python tools/profile_one.py 1024 1 > gpurun_out/gemm_trace.log 2>&1
cp /tmp/product.so $L/libvit_b200.so
grep -c gemm_trace gpurun_out/gemm_trace.log
