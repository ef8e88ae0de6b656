#!/bin/bash
# ncu evidence for the bench command (each capture only after the same command exited 0 without ncu):
#   1. launch list (per-launch device time) of two full steps
#   2. --set full captures of one vision layer's four GEMMs (qkv, out-proj, fc1, fc2) and one vision attention launch
# Towers are serialised (VLMCLIP_OVERLAP_TOWERS=0) so launch indices are deterministic: per step 48 text GEMMs, the patch
# GEMM, then 4 GEMMs per vision layer.
mkdir -p gpurun_out
export VLMCLIP_OVERLAP_TOWERS=0
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 520 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch-list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 360 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_pp -s 40 -c 1 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn rc=$?"
ls -la gpurun_out/*.ncu-rep
