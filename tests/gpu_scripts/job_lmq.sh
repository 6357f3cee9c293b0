#!/bin/bash
# LM-in-rounds kernel: its GPU tests, the config-4 A/B over round budgets, a noise-volume run, ncu --set full of the kernel.
# Usage (under gpurun): bash tests/gpu_scripts/job_lmq.sh <tag>
tag=${1:-x}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rounds or biexp or config4 or degenerate or ybounds or linear" > $out/gputest_lmq_$tag.log 2>&1; echo "pytest rc=$?" >> $out/gputest_lmq_$tag.log
tail -5 $out/gputest_lmq_$tag.log
python tests/gpu_scripts/biexp_c4.py 5 > $out/biexp_c4_$tag.log 2>&1; cp $out/biexp_c4.json $out/biexp_c4_$tag.json 2>/dev/null
tail -16 $out/biexp_c4_$tag.log
ncu --set full --clock-control none --import-source on -k regex:fit_kernel_lmq -s 2 -c 1 -f -o $out/prof_lmq_$tag python tests/gpu_scripts/biexp_c4.py 1 4,3 > $out/ncu_lmq_$tag.log 2>&1
ls -la $out | tail -8
