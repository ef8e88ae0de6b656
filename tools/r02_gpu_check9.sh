#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "finalises" 2>&1 | tail -3
timeout 200 python tools/ab_step.py fused_stats 100 4 2>&1 | grep -v Warn | tail -5 > gpurun_out/r02_ab9.log 2>&1
cat gpurun_out/r02_ab9.log
