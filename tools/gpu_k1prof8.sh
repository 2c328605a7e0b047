#!/bin/bash
out=gpurun_out/${1:-k1prof8}; mkdir -p $out
for th in 256 512; do
echo "== verify mode (PIPE=1), 1/8 of the chunks, k1_threads=$th"
PIPE=1 CFG=2 WORLD=8 XT_OPTS="k1_threads=$th" XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 600 python tools/k1_phase_prof.py 1000000 2>&1 | tail -24
done | tee $out/k1_phase_verify_w8.log
