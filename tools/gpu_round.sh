#!/bin/bash
# One gpurun call: GPU tests, bench (both arms), ncu launch list and one full capture of the top kernels.
# Usage on the box:  bash tools/gpu_round.sh [tag]
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; tail -c 3000 $out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"; cat $out/bench_ref.json
# launch list of the same command (own kernels only; torch's generator kernels are excluded by the name filter)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k1_plan|k2_replay|k3_predict|k_reduce' -c 1200 --csv \
    --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $out/launches_bench.log 2>&1; echo "ncu list rc=$?"
# (XT_BENCH_TWO_PHASE: one plan launch and one replay launch per evaluation, so that -s/-c pick whole-data-set launches)
XT_BENCH_TWO_PHASE=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k2_replay_fused|k1_plan' -s 4 -c 2 \
    -o $out/prof_k12 python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $out/prof_bench.log 2>&1; echo "ncu full rc=$?"
ls -la $out
