#!/bin/bash
# ncu launch list (per-launch device time) of the bench command; the plain run comes first and must exit 0
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 520 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch-list rc=$?"; cut -c1-300 gpurun_out/plain_bench.json
