#!/bin/bash
# ncu --set full of the FP32 replay kernel (one whole-data-set launch of the two-phase mode)
out=gpurun_out/${1:-ncuf32}; mkdir -p $out
XT_BENCH_TWO_PHASE=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k2_replay_f32' -s 2 -c 1 \
    -o $out/prof_f32 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > $out/prof_bench.log 2>&1; echo "ncu full rc=$?"
ls -la $out
