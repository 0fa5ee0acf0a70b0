#!/bin/bash
# developer tool: build sort_lab variants.  usage: build_lab.sh <suffix> "<label>" [-D...]
set -e
cd "$(dirname "$0")"
s=$1; shift; label=$1; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false --expt-relaxed-constexpr \
     -DLAB_NAME="\"$label\"" "$@" sort_lab.cu -o lab_$s 2>&1 | grep -v "warning\|^$\|Remark" || true
