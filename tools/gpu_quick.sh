#!/bin/bash
# One development iteration on the GPU box: parity tests, strong-scaling probe, optional phase cycles.
out=gpurun_out/${1:-q}; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/pytest_gpu.log
timeout 600 python tools/strong_probe.py 1000000 ${2:-1,8} 2>&1 | grep -v "^$" | tee $out/strong_probe.log
