#!/bin/bash
# ncu on the key-range split attention at the config-3 shape: per-launch durations, then one full capture of each launch.
mkdir -p gpurun_out
timeout 35 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attention -c 12 --csv --log-file gpurun_out/split_launches.csv python tools/attn_only.py 512 257 16 > gpurun_out/split_ncu.log 2>&1
tail -4 gpurun_out/split_launches.csv | cut -c1-300
timeout 40 ncu --set full --clock-control none --import-source on -k regex:attention_pp -s 6 -c 2 -f -o gpurun_out/split_attn python tools/attn_only.py 512 257 16 >> gpurun_out/split_ncu.log 2>&1
ls -la gpurun_out/
