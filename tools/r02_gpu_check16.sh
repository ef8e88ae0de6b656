#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "test_gemm" > gpurun_out/r02_pytest16a.log 2>&1; echo "pytest16a rc=$?" >> gpurun_out/r02_pytest16a.log
tail -5 gpurun_out/r02_pytest16a.log
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_model.py -m gpu -q --tb=short -x > gpurun_out/r02_pytest16b.log 2>&1; echo "pytest16b rc=$?" >> gpurun_out/r02_pytest16b.log
tail -5 gpurun_out/r02_pytest16b.log
timeout 200 python bench.py --steps 10 --warmup 3 --workload cfg5 > gpurun_out/r02_cfg5_splitk.json 2> gpurun_out/r02_cfg5_splitk.err; echo "cfg5 splitk rc=$?"; cut -c1-260 gpurun_out/r02_cfg5_splitk.json
VLMCLIP_WGRAD_SPLITK=0 timeout 200 python bench.py --steps 10 --warmup 3 --workload cfg5 > gpurun_out/r02_cfg5_nosplitk.json 2> gpurun_out/r02_cfg5_nosplitk.err; echo "cfg5 nosplitk rc=$?"; cut -c1-260 gpurun_out/r02_cfg5_nosplitk.json
