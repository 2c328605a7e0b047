#!/bin/bash
# parity tests + secondary configs (3a, 5, 3) with the current build
out=gpurun_out/${1:-cfgab}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $out/pytest_gpu.log
for c in 3a 5 3; do
  timeout 420 python tools/bench_configs.py --only $c > $out/config_$c.json 2> $out/config_$c.err; echo "config $c rc=$?"
  cat $out/config_$c.json; tail -2 $out/config_$c.err
done
