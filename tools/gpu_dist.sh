#!/bin/bash
# N-GPU checks on one box: sharded objective / predict parity (NCCL) and the bench at N ranks
N=${1:-2}; out=gpurun_out/${2:-dist}; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | grep -v "^$" | tail -12 | tee $out/dist_check_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 3 2> $out/bench_n$N.err | tail -1 | tee $out/bench_n$N.json
