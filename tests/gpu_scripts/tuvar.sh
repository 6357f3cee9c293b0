#!/bin/bash
# TEST TOOLING: tuvar.sh <name> <translation unit, e.g. inst_biexp_f32_e13_16> [nvcc -D flags...] builds
# dosma_b200/libdfit_<name>.so = the main build's objects with that one translation unit recompiled with the given flags.
# For A/B timing of kernel variants (DOSMA_B200_LIB selects the library).  Needs an up-to-date main build.
set -e
name=$1; tu=$2; shift 2
root=$(cd "$(dirname "$0")/../.." && pwd)
cs=$root/dosma_b200/csrc
b=$cs/_build_$name
mkdir -p $b
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 -Xptxas -v --expt-relaxed-constexpr -I $root/include"
# (several translation units: comma-separated)
objs=$(ls $cs/_build/*.o | grep -v stubs.o)
new=""
for t in ${tu//,/ }; do
  nvcc $F "$@" -c $cs/$t.cu -o $b/$t.o > $b/log_$t.txt 2>&1 || { tail -20 $b/log_$t.txt; exit 1; }
  objs=$(echo "$objs" | grep -v "/$t.o")
  new="$new $b/$t.o"
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $root/dosma_b200/libdfit_$name.so $objs $new
echo "$name: built $(ls -la $root/dosma_b200/libdfit_$name.so | awk '{print $5}') bytes"
