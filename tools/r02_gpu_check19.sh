#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "mn_major or split_reduction" > gpurun_out/r02_pytest19a.log 2>&1; echo "pytest19a rc=$?" >> gpurun_out/r02_pytest19a.log
grep "passed\|failed\|Error\|assert" gpurun_out/r02_pytest19a.log | head -12
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -q --tb=short -x > gpurun_out/r02_pytest19b.log 2>&1; echo "pytest19b rc=$?" >> gpurun_out/r02_pytest19b.log
tail -4 gpurun_out/r02_pytest19b.log
for v in mn splitk; do VLMCLIP_WGRAD=$v timeout 200 python bench.py --steps 10 --warmup 3 --workload cfg5 > gpurun_out/r02_cfg5_$v.json 2> gpurun_out/r02_cfg5_$v.err; echo "cfg5 $v rc=$?"; cut -c1-230 gpurun_out/r02_cfg5_$v.json; done
