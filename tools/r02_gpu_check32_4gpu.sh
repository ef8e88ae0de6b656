#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus 4 --steps 10 --warmup 3 --workload cfg5 > gpurun_out/r02_bench_cfg5_4gpu.json 2> gpurun_out/r02_bench_cfg5_4gpu.err; echo "cfg5 rc=$?"
tail -2 gpurun_out/r02_bench_cfg5_4gpu.err | cut -c1-200; cut -c1-400 gpurun_out/r02_bench_cfg5_4gpu.json
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_cfg5_4gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('collective'))"
