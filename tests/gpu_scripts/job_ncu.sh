#!/bin/bash
# One ncu --set full capture per call.  Usage: job_ncu.sh <tag> mono2|lmq_biexp|lmq_mono
tag=${1:-x}
out=gpurun_out
mkdir -p $out
case $2 in
  mono2) ncu --set full --clock-control none --import-source on -k regex:fit_kernel_mono2_tma -s 3 -c 1 -f -o $out/prof_mono2_$tag python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_f_$tag.log 2>&1 ;;
  lmq_biexp) ncu --set full --clock-control none --import-source on -k regex:fit_kernel_lmq -s 2 -c 1 -f -o $out/prof_lmq_biexp_$tag python tests/gpu_scripts/biexp_c4.py 1 5,2 > $out/ncu_b_$tag.log 2>&1 ;;
  lmq_mono) ncu --set full --clock-control none --import-source on -k regex:fit_kernel_lmq -s 1 -c 1 -f -o $out/prof_lmq_mono_$tag python tests/gpu_scripts/lm_noise.py 2 > $out/ncu_m_$tag.log 2>&1 ;;
esac
ls -la $out | tail -3
