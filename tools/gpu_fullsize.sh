#!/bin/bash
# BASELINE configs 3, 4, 5 at their full sizes on ONE GPU (config 3: 10^6 tracks, config 4: 10^7 tracks, config 5: 10^5 tracks)
out=gpurun_out/${1:-full}; mkdir -p $out
timeout 600 python tools/bench_configs.py --only 3 --scale 10 > $out/config_3_full.json 2> $out/config_3_full.err; echo "config 3 rc=$?"; cat $out/config_3_full.json
timeout 600 python tools/bench_configs.py --only 5 --scale 5 > $out/config_5_full.json 2> $out/config_5_full.err; echo "config 5 rc=$?"; cat $out/config_5_full.json
timeout 900 python tools/bench_configs.py --only 4 --scale 10 > $out/config_4_full.json 2> $out/config_4_full.err; echo "config 4 rc=$?"; cat $out/config_4_full.json; tail -2 $out/config_4_full.err
