#!/bin/bash
# A/B timing of engine builds (extrack_b200/variants/*.so travel with the snapshot, git-ignored)
out=gpurun_out/${1:-tune}; mkdir -p $out
for lib in "" $(ls extrack_b200/variants/*.so 2>/dev/null); do
  echo "=== ${lib:-default}"
  XT_LIB_PATH=$lib timeout 600 python tools/tune_k2.py 2>&1 | grep -v "^$" | tee -a $out/tune.log
done
