#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
echo "== forward S=257: narrow g4 / wide g1"
VLMCLIP_ATTN_MMA_WIDE=0 VLMCLIP_ATTN_SPLIT=0 timeout 60 python tools/attn_only.py 512 257 16
VLMCLIP_ATTN_WIDE_GROUP=1 VLMCLIP_ATTN_SPLIT=0 timeout 60 python tools/attn_only.py 512 257 16
echo "== forward S=197 forced mma: narrow g4 / wide g1"
VLMCLIP_ATTN_MMA_WIDE=0 VLMCLIP_ATTN_FORCE_MMA=1 timeout 60 python tools/attn_only.py 256 197 12
VLMCLIP_ATTN_WIDE_GROUP=1 VLMCLIP_ATTN_FORCE_MMA=1 timeout 60 python tools/attn_only.py 256 197 12
for w in 4 5 6; do
  echo "== bwd variant 2, warps $w"
  VLMCLIP_ATTN_BWD_WARPS=$w timeout 120 python tools/attn_bwd_only.py
done
timeout 200 python tools/ft_bench.py 256 5 2>&1 | grep -a "ms_per_step\|attention" | cut -c1-300
