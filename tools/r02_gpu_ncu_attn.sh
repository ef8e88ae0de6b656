#!/bin/bash
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_bwd -s 6 -c 2 -o gpurun_out/r02_attn_bwd -f python tools/attn_bwd_only.py > gpurun_out/r02_ncu_attn_bwd.log 2>&1; echo "ncu bwd rc=$?"
VLMCLIP_ATTN_SPLIT=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -s 3 -c 1 -o gpurun_out/r02_attn_fwd257 -f python tools/attn_only.py 128 257 16 > gpurun_out/r02_ncu_attn_fwd.log 2>&1; echo "ncu fwd rc=$?"
ls -la gpurun_out/*.ncu-rep
