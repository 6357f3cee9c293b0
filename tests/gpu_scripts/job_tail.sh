#!/bin/bash
# LM tail of the dense kernel: GPU tests, noise volumes with / without the tail and over round budgets, headline A/B.
tag=${1:-x}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tail or rounds or low_snr or degenerate or golden" > $out/gputest_tail_$tag.log 2>&1; echo "pytest rc=$?" >> $out/gputest_tail_$tag.log
tail -4 $out/gputest_tail_$tag.log
for cfg in "1 6,4" "0 6,4" "1 6,6" "1 8,4" "1 4,3" "1 8,8"; do
  set -- $cfg
  DFIT_LM_TAIL=$1 DFIT_LMQ=$2 python tests/gpu_scripts/noise_volume.py > $out/noise_${tag}_$1_$2.log 2>&1
  cp $out/noise_volume.json $out/noise_${tag}_$1_$2.json
  python - <<PY
import json
d = json.load(open("$out/noise_volume.json"))
print("tail=$1 lmq=$2", {k: {m: round(v[m]["ms"], 3) for m in ("default", "lm_only", "tissue_mask")} for k, v in d.items() if isinstance(v, dict)})
PY
done
for t in 1 0; do DFIT_LM_TAIL=$t python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tail=$t', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d.get('gpu_launches'))"; done
