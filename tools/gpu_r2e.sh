#!/bin/bash
out=gpurun_out/${1:-r2e}; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/pytest_gpu.log
timeout 300 python tools/strong_probe.py 1000000 1,2,4,8 2>&1 | grep world | tee $out/strong_probe.log
WORLD=8 XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 300 python tools/k1_phase_prof.py 2>&1 | grep -v "^chunk" | grep -A12 "L=30" | tee $out/k1_phase_w8.log
