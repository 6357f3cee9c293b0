#!/bin/bash
# Round-end single-GPU job: GPU test-suite, bench, every BASELINE configuration, config 4 over the LM kernels,
# noise-dominated volumes, ncu launch list of the bench.  (ncu --set full captures: job_ncu.sh, one kernel per call -- the
# reports are ~40 MB each and gpurun brings back 64 MB.)
tag=${1:-x}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q > $out/gputest_$tag.log 2>&1; echo "pytest rc=$?" >> $out/gputest_$tag.log
tail -4 $out/gputest_$tag.log
python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -c 600 $out/bench_$tag.json; tail -3 $out/bench_$tag.err
python tests/gpu_scripts/configs_bench.py > $out/configs_$tag.json 2> $out/configs_$tag.err; tail -3 $out/configs_$tag.err
python tests/gpu_scripts/biexp_c4.py 5 5,2 > $out/biexp_c4_$tag.log 2>&1; cp $out/biexp_c4.json $out/biexp_c4_$tag.json; cut -c1-100 $out/biexp_c4_$tag.log
python tests/gpu_scripts/noise_volume.py > $out/noise_$tag.log 2>&1; cp $out/noise_volume.json $out/noise_volume_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_l_$tag.log 2>&1
ls -la $out | tail -5
