#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short > gpurun_out/model.log 2>&1
echo "model rc=$?" >> gpurun_out/model.log
tail -80 gpurun_out/model.log
