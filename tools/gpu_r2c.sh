#!/bin/bash
# bench (new strong / api / secondary blocks) on one GPU + plan-kernel phase cycles for the world-8 shard (512 threads).
out=gpurun_out/${1:-r2c}; mkdir -p $out
timeout 900 python bench.py --steps 50 --no-cpu > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; tail -c 6000 $out/bench.json; tail -5 $out/bench.err
WORLD=8 XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 300 python tools/k1_phase_prof.py 2>&1 | grep -v "^chunk" | tee $out/k1_phase_w8.log
XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 300 python tools/k1_phase_prof.py 2>&1 | grep -v "^chunk" | grep -A12 "L=30" | tee $out/k1_phase_w1.log
