#!/bin/bash
# Full GPU test suite + smoke + one bench line (no profiler)
out=gpurun_out/${1:-check}; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log; tail -5 $out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; tail -c 4000 $out/bench.json; tail -5 $out/bench.err
