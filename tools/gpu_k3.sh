#!/bin/bash
out=gpurun_out/${1:-k3}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "predict or refine or nbmax or window" 2>&1 | tail -5
{
for lib in "" extrack_b200/variants/libxt_k3c3.so extrack_b200/variants/libxt_k3c4.so; do
 for c in 40 48 64; do
  echo "== lib=${lib:-default} cap0=$c"
  XT_LIB_PATH=$lib XT_OPTS="k3_cap0=$c" timeout 600 python tools/bench_configs.py --only 4 2>&1 | tail -1
 done
done
} | tee $out/k3_ab.log
