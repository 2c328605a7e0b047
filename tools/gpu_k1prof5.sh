#!/bin/bash
out=gpurun_out/${1:-k1prof5}; mkdir -p $out
CFG=5 XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 600 python tools/k1_phase_prof.py 20000 > $out/k1_phase_cfg5.log 2>&1; echo rc=$?
cat $out/k1_phase_cfg5.log
