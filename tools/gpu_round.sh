#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (both arms), ncu launch list and full captures of the top kernels.
# Usage on the box:  bash tools/gpu_round.sh [tag]
tag=${1:-r12}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; tail -c 1500 $out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"; cat $out/bench_ref.json
timeout 300 python tools/stress_determinism.py 200 > $out/stress_determinism.log 2>&1; tail -1 $out/stress_determinism.log
# launch list of the same command (own kernels only; torch's generator kernels are excluded by the name filter)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k1_plan|k2_replay|k3_|k_reduce|k_pack|k_refine' -c 1500 --csv \
    --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary --e2e-steps 1 > $out/launches_bench.log 2>&1; echo "ncu list rc=$?"
# construction mode (XT_BENCH_TWO_PHASE: one plan launch and one replay launch per evaluation): whole-data-set launches
XT_BENCH_TWO_PHASE=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k2_replay_fused|k1_plan' -s 4 -c 2 \
    -o $out/prof_k12 python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary --e2e-steps 1 > $out/prof_bench.log 2>&1; echo "ncu full (construction) rc=$?"
# default mode: the verification launch of the plan kernel and the replay launch next to it
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k2_replay_fused|k1_plan' -s 8 -c 2 \
    -o $out/prof_verify python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary --e2e-steps 1 > $out/prof_verify.log 2>&1; echo "ncu full (verification) rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_predict -c 1 -o $out/prof_k3 \
    python tools/bench_configs.py --only 4 --scale 0.2 > $out/prof_k3.log 2>&1; echo "ncu k3 rc=$?"
ls -la $out
