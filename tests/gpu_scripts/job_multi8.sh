#!/bin/bash
# 8-GPU job: bench.py at N = 8 and 4 with peer stores (default; `all`: also N = 8 with NVLS multicast stores), configs 3 / 5 at N = 8 and N = 4.
tag=${1:-x}
out=gpurun_out
mkdir -p $out
run_bench() {  # n, multicast mode
  DFIT_BENCH_MULTICAST=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $1 --steps 20 --warmup 5 > $out/bench_${tag}_$1gpu_$2.json 2> $out/bench_${tag}_$1gpu_$2.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/bench_${tag}_$1gpu_$2.json").read().strip().splitlines()[-1])
    print("N=$1 $2:", d["gather"], "| ms/step", round(d["ms_per_step"],4), "value %.4g" % d["value"], "nvlink GB/s", round(d["roofline"]["achieved"],1), "sustained ms", round(d["sustained"]["ms_per_step"],4), "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print("N=$1 $2 failed:", e); print(open("$out/bench_${tag}_$1gpu_$2.err").read()[-1500:])
PY
}
run_bench 8 off
[ "$2" = all ] && run_bench 8 auto
run_bench 4 off
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 tests/gpu_scripts/configs_multi.py > $out/configs_multi_${tag}_${n}gpu.json 2> $out/configs_multi_${tag}_${n}gpu.err
tail -40 $out/configs_multi_${tag}_${n}gpu.json; tail -3 $out/configs_multi_${tag}_${n}gpu.err
done
