#!/bin/bash
# usage: scripts/build_variant.sh "<extra nvcc flags>" [name]
#   rebuilds the library with extra -D flags.  With a name the result goes to mahakala_b200/variants/lib<name>.so
#   (select it with MAHAKALA_B200_LIB=...) and the default library is rebuilt afterwards.
set -e
cd "$(dirname "$0")/../mahakala_b200/csrc"
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr -ccbin /usr/bin/g++"
make clean >/dev/null
make -j8 NVCCFLAGS="$FLAGS $1" 2>&1 | grep -E "error" || true
if [ -n "$2" ]; then
  mkdir -p ../variants
  cp ../libmahakala_b200.so ../variants/lib$2.so
  grep -A2 "render_kernelILi[18]E\|integrate_kernelINS_10KerrSchildELi2" render.o.ptxas.log integrate.o.ptxas.log | grep -i "registers\|spill" | tr "\n" " "; echo
  make clean >/dev/null
  make -j8 2>&1 | grep -E "error" || true
fi
ls -la ../libmahakala_b200.so | awk '{print $5}'
