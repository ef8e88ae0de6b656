#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 500 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "attention" 2>&1 | tail -4
for w in 0 1 0 1; do VLMCLIP_ATTN_SPLIT=0 VLMCLIP_ATTN_MMA_WIDE=$w timeout 60 python tools/attn_only.py 512 257 16; done
VLMCLIP_ATTN_SPLIT=3 timeout 60 python tools/attn_only.py 512 257 16
