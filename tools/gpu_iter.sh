#!/bin/bash
# One development iteration on the GPU box: parity tests, plan-kernel phase cycles (profiling build in
# extrack_b200/variants/libxt_prof.so if present), replay / pipeline timing.
out=gpurun_out/${1:-iter}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/pytest_gpu.log
if [ -f extrack_b200/variants/libxt_prof.so ]; then
  XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 300 python tools/k1_phase_prof.py 2>&1 | tee $out/k1_phase.log
fi
timeout 600 python tools/tune_k2.py 2>&1 | grep -v "^$" | tee $out/tune.log
