#!/bin/bash
# Source-level ncu capture of the plan kernel on the world-8 shard (64 chunks, 512 threads per chunk).
out=gpurun_out/${1:-ncu_k1w8}; mkdir -p $out
XT_OPTS="pipeline=0" timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_plan -s 4 -c 1 -o $out/prof_k1w8 \
  python tools/strong_probe.py 1000000 8 > $out/probe.log 2>&1; echo "ncu rc=$?"; tail -3 $out/probe.log
