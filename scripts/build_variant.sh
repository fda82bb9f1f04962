#!/bin/bash
# usage: scripts/build_variant.sh "<extra nvcc flags>"  -> rebuilds the library with extra -D flags
cd "$(dirname "$0")/../mahakala_b200/csrc" && make clean >/dev/null && make -j8 NVCCFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr -ccbin /usr/bin/g++ $1" 2>&1 | grep -E "error" ; ls -la ../libmahakala_b200.so | awk '{print $5}'
