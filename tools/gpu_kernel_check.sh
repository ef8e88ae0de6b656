#!/bin/bash
# First-contact check of every kernel on a B200: non-GEMM tests first, then the tcgen05 GEMM under its own
# timeout so a hang there cannot eat the other results.  Output goes to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "not gemm" > gpurun_out/kernels_other.log 2>&1
echo "other rc=$?" >> gpurun_out/kernels_other.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "gemm" > gpurun_out/kernels_gemm.log 2>&1
echo "gemm rc=$?" >> gpurun_out/kernels_gemm.log
tail -40 gpurun_out/kernels_other.log
tail -60 gpurun_out/kernels_gemm.log
