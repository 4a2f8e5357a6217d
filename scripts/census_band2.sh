#!/bin/bash
# hot-path instruction count of warp_bwd_band2 (slow path and clamped scatter compiled out)
cd "$(dirname "$0")/../dssmd_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -I ../../include -DBAND2_CENSUS -c warp_bwd_band2.cu -o /tmp/sass/band2c.o && cd ../.. && python scripts/sass_census.py /tmp/sass/band2c.o "${1:-band2_kernelILi3ELi258ELb1}" $2
