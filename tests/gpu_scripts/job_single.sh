#!/bin/bash
# Single-GPU job: GPU test-suite, bench (with the Python-API legs), noise-dominated volume, configs.
tag=${1:-x}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q > $out/gputest_$tag.log 2>&1; echo "pytest rc=$?" >> $out/gputest_$tag.log
tail -5 $out/gputest_$tag.log
python bench.py --steps 20 --warmup 5 --no-cpu > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -c 2800 $out/bench_$tag.json; tail -3 $out/bench_$tag.err
python tests/gpu_scripts/noise_volume.py > $out/noise_$tag.log 2>&1; cp $out/noise_volume.json $out/noise_volume_$tag.json; tail -5 $out/noise_$tag.log | cut -c1-900
python tests/gpu_scripts/configs_bench.py > $out/configs_$tag.json 2> $out/configs_$tag.err
