#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
SECONDS=0
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$? wall ${SECONDS}s"
tail -2 gpurun_out/r02_final_bench.err | cut -c1-200; cat gpurun_out/r02_final_bench.json
