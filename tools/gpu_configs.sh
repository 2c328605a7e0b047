#!/bin/bash
# One gpurun call: secondary configs (3a, 3, 4, 5) at bounded sizes, each under its own timeout.
tag=${1:-cfg}
out=gpurun_out/$tag
mkdir -p $out
for c in 3a 4 5 3; do
  timeout 420 python tools/bench_configs.py --only $c > $out/config_$c.json 2> $out/config_$c.err; echo "config $c rc=$?"
  cat $out/config_$c.json; tail -3 $out/config_$c.err
done
