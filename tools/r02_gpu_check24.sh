#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "variants or key_range" 2>&1 | tail -4
for v in 3 4; do VLMCLIP_ATTN_SPLIT=$v timeout 60 python tools/attn_only.py 512 257 16; done
