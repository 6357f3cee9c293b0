#!/bin/bash
# TEST TOOLING: quickvar.sh <name> [nvcc -D flags...] builds dosma_b200/libdfit_<name>.so in which only the 1..8-echo fp32
# mono-exponential kernels exist (compiled with the given flags); every other model / dtype returns "not supported".
# For A/B timing of the headline kernel with tests/gpu_scripts/ab_libs.py.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/../.." && pwd)
cs=$root/dosma_b200/csrc
b=$cs/_build_$name
mkdir -p $b
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC,-O2 -Xptxas -v --expt-relaxed-constexpr -I $root/include"
nvcc $F "$@" -c $cs/inst_mono_f32_lo.cu -o $b/inst_mono_f32_lo.o > $b/log.txt 2>&1 || { tail -20 $b/log.txt; exit 1; }
[ -f $cs/_build/stubs.o ] || nvcc $F -c $root/tests/gpu_scripts/variant_stubs/inst_stubs.cu -o $cs/_build/stubs.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $root/dosma_b200/libdfit_$name.so $cs/_build/dfit_api.o $cs/_build/qdess.o $cs/_build/metrics.o $cs/_build/stubs.o $b/inst_mono_f32_lo.o
echo "$name: $(grep -A3 'fit_kernel_mono2_tmaINS_7MonoExpELi8ELb0Ef' $b/log.txt | grep -i 'used' | head -1) $(ls -la $root/dosma_b200/libdfit_$name.so | awk '{print $5}') bytes"
