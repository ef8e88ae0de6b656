#!/bin/bash
# Round 2, fourth GPU pass: strip loss kernels, ring prefetcher, full GPU suite, bench K=20.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -s -k "clip_loss or class_head" > gpurun_out/r02_pytest4a.log 2>&1; echo "pytest4a rc=$?" >> gpurun_out/r02_pytest4a.log
tail -25 gpurun_out/r02_pytest4a.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest4.log
tail -15 gpurun_out/r02_pytest4.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-full-finetune > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench4.err; cat gpurun_out/r02_bench4.json
