#!/bin/bash
# developer tool: build a libusrt variant with extra -D switches for trace.cu only.  usage: build_trace_lab.sh <suffix> [-D...]
set -e
cd "$(dirname "$0")/../.."
s=$1; shift
B=unitysimpleraytracing_b200/csrc/build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
     -c unitysimpleraytracing_b200/csrc/trace.cu -o /tmp/trace_$s.o
nvcc -shared -o tools/micro/libusrt_$s.so $B/api.o $B/morton.o $B/radix_sort.o $B/lbvh.o /tmp/trace_$s.o $B/shade.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
cuobjdump -res-usage tools/micro/libusrt_$s.so 2>/dev/null | grep -A1 "k_trace_primaryILb0" | grep -o "REG:[0-9]* STACK:[0-9]* SHARED:[0-9]*" | head -1
