#!/bin/bash
# One GPU call of the kernel work loop: GPU test-suite, kernel variants, a short bench, launch list + ncu --set full of
# the headline kernel.  Usage (under gpurun): bash tests/gpu_scripts/job_kernel.sh <tag>
tag=${1:-x}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/gputest_$tag.log 2>&1; echo "pytest rc=$?" >> $out/gputest_$tag.log
tail -5 $out/gputest_$tag.log
python tests/gpu_scripts/variants.py > $out/variants_$tag.log 2>&1; cp $out/variants.json $out/variants_$tag.json 2>/dev/null
tail -6 $out/variants_$tag.log
python bench.py --steps 20 --warmup 5 --no-cpu > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -c 1500 $out/bench_$tag.json
python tests/gpu_scripts/configs_bench.py > $out/configs_$tag.json 2> $out/configs_$tag.err; tail -c 3000 $out/configs_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fit_kernel_mono2_tma -s 3 -c 1 -f -o $out/prof_$tag python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_f_$tag.log 2>&1
ls -la $out | tail -12
