#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "attention" 2>&1 | tail -3
VLMCLIP_ATTN_FORCE_MMA=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "test_attention and not variants" 2>&1 | tail -3
for v in 0 1; do
  VLMCLIP_ATTN_FORCE_MMA=$v timeout 60 python tools/attn_only.py 256 197 12
  VLMCLIP_ATTN_FORCE_MMA=$v timeout 60 python tools/attn_only.py 512 257 16
done
timeout 60 python tools/kernel_bench.py 2>&1 | grep "attention"
