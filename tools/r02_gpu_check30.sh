#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
for w in 8 5 4; do echo "== forward S=257, max warps $w"; VLMCLIP_ATTN_FWD_WARPS=$w timeout 60 python tools/attn_only.py 512 257 16; done
for w in 8 4; do echo "== forward S=197 forced mma, max warps $w"; VLMCLIP_ATTN_FWD_WARPS=$w VLMCLIP_ATTN_FORCE_MMA=1 timeout 60 python tools/attn_only.py 256 197 12; done
echo "== bwd variant 2, warps 3"
VLMCLIP_ATTN_BWD_WARPS=3 timeout 120 python tools/attn_bwd_only.py
