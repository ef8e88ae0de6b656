#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_fullsize.py -m gpu -q --tb=short -x -k "attention or fullsize or L14 or l14 or large" 2>&1 | tail -4
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 > gpurun_out/r02_cfg3_wide.json 2> gpurun_out/r02_cfg3_wide.err; echo "cfg3 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02_cfg3_wide.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"
