#!/bin/bash
# Round 2, third GPU pass: model tests with the two-graph pipelined step, bench at K=20 and K=100.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_golden.py -m gpu -q --tb=short -x > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest3.log
tail -15 gpurun_out/r02_pytest3.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-full-finetune > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_bench3.err; cat gpurun_out/r02_bench3.json
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-full-finetune > gpurun_out/r02_bench3_100.json 2> gpurun_out/r02_bench3_100.err; echo "bench100 rc=$?"
cat gpurun_out/r02_bench3_100.json
