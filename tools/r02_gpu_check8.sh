#!/bin/bash
# Round 2, eighth GPU pass (1 GPU): whole suite with the fused statistics, interleaved A/Bs.
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest8.log
tail -8 gpurun_out/r02_pytest8.log
for w in fused_stats residual graph; do timeout 200 python tools/ab_step.py $w 100 4 2>&1 | grep -v Warn | tail -5; done > gpurun_out/r02_ab8.log 2>&1
cat gpurun_out/r02_ab8.log
