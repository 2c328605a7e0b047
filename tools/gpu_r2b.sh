#!/bin/bash
# Plan-kernel phase cycles for the world-8 shard (1024 and 256 threads per chunk) and one source-level ncu capture of
# the two-phase plan / replay launches.
out=gpurun_out/${1:-r2b}; mkdir -p $out
WORLD=8 XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 300 python tools/k1_phase_prof.py 2>&1 | grep -v "^chunk" | tee $out/k1_phase_w8.log
WORLD=8 XT_OPTS="k1_threads=256" XT_LIB_PATH=extrack_b200/variants/libxt_prof.so timeout 300 python tools/k1_phase_prof.py 2>&1 | grep -v "^chunk" | tee $out/k1_phase_w8_256.log
XT_BENCH_TWO_PHASE=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k2_replay_fused|k1_plan' -s 4 -c 2 \
    -o $out/prof_k12 python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $out/prof_bench.log 2>&1; echo "ncu full rc=$?"
ls -la $out
