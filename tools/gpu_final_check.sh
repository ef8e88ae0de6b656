#!/bin/bash
# Round-end pass on one B200: every GPU test, smoke(), the config-3 bench (ViT-L/14, the shape the key-range split serves).
mkdir -p gpurun_out
timeout 110 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 40 python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_1gpu.json 2> gpurun_out/bench_cfg3_1gpu.err; echo "cfg3 rc=$?"
cut -c1-330 gpurun_out/bench_cfg3_1gpu.json
