#!/bin/bash
# Multi-GPU job: the gather check (all transports) and bench.py --gpus N with the fused peer stores and the copy-engine pipeline.
tag=${1:-x}; n=${2:-2}
out=gpurun_out
mkdir -p $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/gpu_scripts/check_fused_gather.py 200192 200068 > $out/check_$tag.log 2>&1
grep -c "True" $out/check_$tag.log; grep -n "MISMATCH\|False\|Error" $out/check_$tag.log | cut -c1-300 | head
for mode in stores copies; do
DFIT_BENCH_GATHER=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 20 --warmup 5 > $out/bench_${tag}_${n}gpu_$mode.json 2> $out/bench_${tag}_${n}gpu_$mode.err
python - <<PY
import json
try:
    d=json.loads(open("$out/bench_${tag}_${n}gpu_$mode.json").read().strip().splitlines()[-1])
    print("$mode", d["gather"], "ms/step", round(d["ms_per_step"],4), "value", d["value"], "nvlink", d["roofline"]["achieved"], "sustained ms", d["sustained"]["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
except Exception as e:
    print("$mode FAILED", e); print(open("$out/bench_${tag}_${n}gpu_$mode.err").read()[-1500:])
PY
done
