#!/bin/bash
out=gpurun_out/${1:-k2h}; mkdir -p $out
{
for h in 0 1 2 3; do echo "== warp-0 handicap $h"; XT_OPTS="k2_cost_w0=$h" timeout 300 python tools/strong_probe.py 1000000 1 2>&1 | grep world; done
echo "== no x2 unroll, handicap 1"; XT_LIB_PATH=extrack_b200/variants/libxt_nox2.so timeout 300 python tools/strong_probe.py 1000000 1 2>&1 | grep world
echo "== default, worlds"; timeout 300 python tools/strong_probe.py 1000000 1,2,4,8 2>&1 | grep world
} | tee $out/k2h.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
