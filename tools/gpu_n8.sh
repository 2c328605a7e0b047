#!/bin/bash
# 8-GPU box: bench.py under torchrun (weak headline + strong / secondary blocks) and the in-process multi-GPU probe.
out=gpurun_out/${1:-n8}; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 100 --warmup 3 > $out/bench_n8.json 2> $out/bench_n8.err; echo "bench rc=$?"
tail -c 4000 $out/bench_n8.json; tail -3 $out/bench_n8.err
timeout 600 python tools/multi_probe.py 2>&1 | tee $out/multi_probe.log
timeout 300 python -m pytest tests -m gpu -x -q -k "two_real_devices or multi_engine" 2>&1 | tail -3 | tee $out/pytest_multi.log
