#!/bin/bash
# Rebuild the engine with per-phase cycle counters in the plan kernel (on the box only) and print them.
out=gpurun_out/${1:-k1prof}; mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DXT_K1_PROF -Iinclude -Iextrack_b200/csrc -shared -Xcompiler -fPIC \
  -o extrack_b200/libxtrack_b200.so extrack_b200/csrc/xt_engine.cu || exit 1
timeout 600 python tools/k1_phase_prof.py 1000000 > $out/k1_phase.log 2>&1; echo rc=$?
cat $out/k1_phase.log
