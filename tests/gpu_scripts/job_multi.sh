#!/bin/bash
# Multi-GPU job: the world-size-2 GPU test, then bench.py at N = the box's GPU count (and N = 1).  Usage: job_multi.sh <tag> <ngpus>
tag=${1:-x}; n=${2:-2}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_multi.py -q -x > $out/gputest_multi_$tag.log 2>&1; echo "pytest rc=$?" >> $out/gputest_multi_$tag.log
tail -4 $out/gputest_multi_$tag.log
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 20 --warmup 5 > $out/bench_${tag}_${n}gpu.json 2> $out/bench_${tag}_${n}gpu.err
tail -c 2500 $out/bench_${tag}_${n}gpu.json; tail -5 $out/bench_${tag}_${n}gpu.err
python bench.py --steps 20 --warmup 5 --no-cpu > $out/bench_${tag}_1gpu.json 2> $out/bench_${tag}_1gpu.err
tail -c 2500 $out/bench_${tag}_1gpu.json; tail -3 $out/bench_${tag}_1gpu.err
