#!/bin/bash
# Parity tests, then sweeps on the config-2 data set: replay-schedule cost model (world 1), plan-kernel threads (world 8 shard).
out=gpurun_out/${1:-sweep}; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/pytest_gpu.log
{
echo "== default"; timeout 300 python tools/strong_probe.py 1000000 1,2,4,8 2>&1 | grep world
for c in "4,6,6,2" "4,6,4,2" "4,6,5,1" "4,5,4,1" "4,7,8,3" "8,13,10,3"; do
  IFS=, read a b cc d <<< "$c"
  echo "== cost model $c"; XT_OPTS="k2_cost0=$a,k2_cost1=$b,k2_cost2=$cc,k2_cost3=$d" timeout 300 python tools/strong_probe.py 1000000 1 2>&1 | grep world
done
for t in 256 512 1024; do
  echo "== world-8 shard, plan kernel with $t threads"; XT_OPTS="k1_threads=$t" timeout 300 python tools/strong_probe.py 1000000 8 2>&1 | grep world
done
} | tee $out/sweep.log
