#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 3 --workload cfg5 > gpurun_out/r02_bench_cfg5_8gpu.json 2> gpurun_out/r02_bench_cfg5_8gpu.err; echo "cfg5 rc=$?"
tail -2 gpurun_out/r02_bench_cfg5_8gpu.err | cut -c1-200; cat gpurun_out/r02_bench_cfg5_8gpu.json
