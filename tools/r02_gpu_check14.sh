#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -s -k "clip_loss" > gpurun_out/r02_pytest14a.log 2>&1; echo "pytest14a rc=$?" >> gpurun_out/r02_pytest14a.log
grep "clip loss strips\|passed\|failed\|Error\|assert" gpurun_out/r02_pytest14a.log | head -12
timeout 100 python tools/loss_only.py > gpurun_out/loss_only.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:clip_ -s 16 -c 3 -o gpurun_out/prof_loss -f python tools/loss_only.py > gpurun_out/ncu_loss.log 2>&1
echo "loss ncu rc=$?"
