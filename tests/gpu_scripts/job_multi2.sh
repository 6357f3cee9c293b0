#!/bin/bash
tag=${1:-x}; n=${2:-2}
out=gpurun_out
mkdir -p $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/gpu_scripts/check_fused_gather.py 200192 > $out/check_$tag.log 2>&1
grep -n "split list\|MISMATCH" $out/check_$tag.log | cut -c1-700
for mc in auto off; do
DFIT_BENCH_MULTICAST=$mc python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 20 --warmup 5 > $out/bench_${tag}_${n}gpu_$mc.json 2> $out/bench_${tag}_${n}gpu_$mc.err
python - <<PY
import json
d=json.loads(open("$out/bench_${tag}_${n}gpu_$mc.json").read().strip().splitlines()[-1])
print("$mc", d["gather"], "ms/step", round(d["ms_per_step"],4), "value", d["value"], "nvlink", d["roofline"]["achieved"], "sustained ms", d["sustained"]["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
