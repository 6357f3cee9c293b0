#!/bin/bash
# GPU test-suite on the default library, A/B timing of every library variant in dosma_b200/, configs bench.
tag=${1:-x}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q > $out/gputest_$tag.log 2>&1; echo "pytest rc=$?" >> $out/gputest_$tag.log
tail -6 $out/gputest_$tag.log
python tests/gpu_scripts/ab_libs.py > $out/ab_$tag.log 2>&1; cp $out/ab_libs.json $out/ab_libs_$tag.json
cat $out/ab_$tag.log
python tests/gpu_scripts/configs_bench.py > $out/configs_$tag.json 2> $out/configs_$tag.err; python - <<PY
import json
d=json.load(open("$out/configs_$tag.json"))
for k,v in d.items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a.startswith("ms") or "per_s" in a})
PY
