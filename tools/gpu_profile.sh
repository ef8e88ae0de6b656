#!/bin/bash
# ncu evidence for the bench command: launch list (per-launch device time) + full captures of the top kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 420 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch-list rc=$?"
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 40 -c 6 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 4 -c 2 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn rc=$?"
ls -la gpurun_out/
