#!/bin/bash
# Round 2, sixth GPU pass (1 GPU): attention with the split S issue, PE-CLIP adapter backward, whole suite, bench.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "attention" > gpurun_out/r02_pytest6a.log 2>&1; echo "pytest6a rc=$?" >> gpurun_out/r02_pytest6a.log
tail -8 gpurun_out/r02_pytest6a.log
for v in 1 0; do
  VLMCLIP_ATTN_SSPLIT=$v timeout 60 python tools/attn_only.py 256 197 12
  VLMCLIP_ATTN_SSPLIT=$v timeout 60 python tools/attn_only.py 512 257 16
  VLMCLIP_ATTN_SSPLIT=$v timeout 60 python tools/attn_only.py 64 224 12
done > gpurun_out/r02_attn_ab.log 2>&1
VLMCLIP_ATTN_DEBUG=1 timeout 60 python tools/attn_only.py 256 197 12 2>&1 | grep "attn-pp dbg" | tail -4 >> gpurun_out/r02_attn_ab.log
VLMCLIP_ATTN_SSPLIT=0 VLMCLIP_ATTN_DEBUG=1 timeout 60 python tools/attn_only.py 256 197 12 2>&1 | grep "attn-pp dbg" | tail -4 >> gpurun_out/r02_attn_ab.log
cat gpurun_out/r02_attn_ab.log
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest6.log
tail -15 gpurun_out/r02_pytest6.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-full-finetune > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench6.err; cat gpurun_out/r02_bench6.json
