#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 51 -c 2 -o gpurun_out/prof_gemm_fc1 $CMD > gpurun_out/ncu_gemm_fc1.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_gemm_fc1.log
