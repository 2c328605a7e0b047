#!/bin/bash
# Round-2 first measurement: strong-scaling probe (one GPU, shards of a world of 1/2/4/8) and plan-kernel phase cycles.
out=gpurun_out/${1:-r2a}; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 600 python tools/strong_probe.py 1000000 1,2,4,8 2>&1 | grep -v "^$" | tee $out/strong_probe.log
XT_OPTS="k1_threads=256" timeout 300 python tools/strong_probe.py 1000000 8 2>&1 | grep world | tee $out/strong_probe_k1_256.log
XT_OPTS="n_groups=1" timeout 300 python tools/strong_probe.py 1000000 1,8 2>&1 | grep world | tee $out/strong_probe_g1.log
XT_OPTS="n_groups=12,n_streams=12" timeout 300 python tools/strong_probe.py 1000000 1,8 2>&1 | grep world | tee $out/strong_probe_g12.log
if [ -f extrack_b200/variants/libxt_prof.so ]; then
  XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 300 python tools/k1_phase_prof.py 2>&1 | tee $out/k1_phase.log
fi
if [ -f extrack_b200/variants/libxt_prof.so ]; then
  WORLD=8 XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 300 python tools/k1_phase_prof.py 2>&1 | tee $out/k1_phase_w8.log
  WORLD=8 XT_OPTS="k1_threads=256" XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 300 python tools/k1_phase_prof.py 2>&1 | tee $out/k1_phase_w8_256.log
fi
