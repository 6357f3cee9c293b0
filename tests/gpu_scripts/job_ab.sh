#!/bin/bash
# GPU test-suite on the default library, then A/B timing of every library variant in dosma_b200/.
tag=${1:-x}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/gputest_$tag.log 2>&1; echo "pytest rc=$?" >> $out/gputest_$tag.log
tail -4 $out/gputest_$tag.log
python tests/gpu_scripts/ab_libs.py > $out/ab_$tag.log 2>&1; cp $out/ab_libs.json $out/ab_libs_$tag.json
cat $out/ab_$tag.log
