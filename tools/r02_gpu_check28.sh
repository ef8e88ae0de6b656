#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 500 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "attention" 2>&1 | tail -3
for v in 0 1 2; do
  echo "== VLMCLIP_ATTN_BWD=$v"
  VLMCLIP_ATTN_BWD=$v timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q --tb=short -x -k "attention" 2>&1 | tail -2
  VLMCLIP_ATTN_BWD=$v timeout 120 python tools/attn_bwd_only.py
done
for g in 2 4; do echo "== wide group $g"; VLMCLIP_ATTN_WIDE_GROUP=$g VLMCLIP_ATTN_SPLIT=0 timeout 60 python tools/attn_only.py 512 257 16; done
VLMCLIP_ATTN_FORCE_MMA=1 timeout 60 python tools/attn_only.py 256 197 12
timeout 200 python tools/ft_bench.py 256 5 2>&1 | tail -3 | cut -c1-400
