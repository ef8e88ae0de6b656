#!/bin/bash
# Round 2, 8-GPU pass: BASELINE configs 2, 3 (global batch 4096) and 5 where BASELINE puts them.
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 --workload cfg3 --lean > gpurun_out/r02_bench_cfg3_8gpu.json 2> gpurun_out/r02_bench_cfg3_8gpu.err; echo "cfg3 rc=$?"
tail -2 gpurun_out/r02_bench_cfg3_8gpu.err | cut -c1-300; cat gpurun_out/r02_bench_cfg3_8gpu.json
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 --workload cfg5 > gpurun_out/r02_bench_cfg5_8gpu.json 2> gpurun_out/r02_bench_cfg5_8gpu.err; echo "cfg5 rc=$?"
tail -2 gpurun_out/r02_bench_cfg5_8gpu.err | cut -c1-300; cat gpurun_out/r02_bench_cfg5_8gpu.json
timeout 240 $TR bench.py --gpus 8 --steps 50 --warmup 3 --lean > gpurun_out/r02_bench_cfg2_8gpu.json 2> gpurun_out/r02_bench_cfg2_8gpu.err; echo "cfg2 rc=$?"
tail -2 gpurun_out/r02_bench_cfg2_8gpu.err | cut -c1-300; cat gpurun_out/r02_bench_cfg2_8gpu.json
