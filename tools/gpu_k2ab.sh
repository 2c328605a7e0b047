#!/bin/bash
# A/B of replay-kernel builds / schedules on the config-2 data set (one GPU).
out=gpurun_out/${1:-k2ab}; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/pytest_gpu.log
{
echo "== default (singles x2 interleaved, 24 warps, LPT schedule)"; timeout 300 python tools/strong_probe.py 1000000 1,8 2>&1 | grep world
echo "== default, round-robin schedule"; XT_OPTS="k2_lpt=0" timeout 300 python tools/strong_probe.py 1000000 1 2>&1 | grep world
for v in w20 nox2; do
  echo "== variant $v"; XT_LIB_PATH=extrack_b200/variants/libxt_$v.so timeout 300 python tools/strong_probe.py 1000000 1 2>&1 | grep world
done
} | tee $out/k2ab.log
