#!/bin/sh
# builds the standalone tuning harnesses (not part of the product library)
set -e
cd "$(dirname "$0")"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ sort_bench.cu -o sort_bench
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ nvlink_probe.cu -o nvlink_probe
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ scatter_tma_probe.cu -o scatter_tma_probe
