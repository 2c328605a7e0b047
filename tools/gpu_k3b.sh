#!/bin/bash
out=gpurun_out/${1:-k3b}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/pytest_gpu.log
{
for pc in 1 4 6 10; do
  echo "== k3_pieces=$pc"
  XT_OPTS="k3_pieces=$pc" timeout 600 python tools/bench_configs.py --only 4 2>&1 | tail -1
done
} | tee $out/k3_pieces.log
XT_OPTS="k3_pieces=1" timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_predict -c 1 -o $out/prof_k3 \
    python tools/bench_configs.py --only 4 --scale 0.2 > $out/prof_k3.log 2>&1; echo "ncu k3 rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json,sys
b = json.load(open(sys.argv[1] if len(sys.argv)>1 else "gpurun_out/k3b/bench.json"))
print(json.dumps(b["secondary"]["config_4"]))
print(b["value"], b["ms_per_step"], b["roofline"]["frac"], b["e2e"]["value"])
PY
