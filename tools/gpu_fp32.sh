#!/bin/bash
# FP32 replay: parity tests of the new path, then A/B timing against the FP64 kernel on config 2
out=gpurun_out/${1:-fp32}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "fp32 or precision" 2>&1 | tail -15 | tee $out/pytest_fp32.log
for lib in "" $(ls extrack_b200/variants/*.so 2>/dev/null); do
  echo "=== ${lib:-default}"
  XT_LIB_PATH=$lib TUNE_FP32=1 TUNE_QUICK=1 timeout 600 python tools/tune_k2.py 2>&1 | grep -v "^$" | tee -a $out/tune.log
done
