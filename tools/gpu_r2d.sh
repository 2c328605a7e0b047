#!/bin/bash
out=gpurun_out/${1:-r2d}; mkdir -p $out
WORLD=8 XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 300 python tools/k1_phase_prof.py 2>&1 | grep -v "^chunk" | tee $out/k1_phase_w8.log
XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 300 python tools/k1_phase_prof.py 2>&1 | grep -v "^chunk" | grep -A12 "L=30" | tee $out/k1_phase_w1.log
