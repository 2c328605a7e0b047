#!/bin/bash
out=gpurun_out/${1:-tail}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/stress_determinism.py 300 2>&1 | tail -3 | tee $out/stress.log
timeout 300 python tools/strong_probe.py 1000000 1,2,4,8 2>&1 | grep world | tee $out/strong.log
