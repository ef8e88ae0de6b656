#!/bin/bash
# Key-range split attention (S = 257): parity tests, A/B of the variants against the mma.sync kernel, phase timers.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "attention" > gpurun_out/split_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/split_pytest.log
tail -8 gpurun_out/split_pytest.log
: > gpurun_out/split_ab.log
for v in 3 2 1; do
  echo "VLMCLIP_ATTN_SPLIT=$v" >> gpurun_out/split_ab.log
  VLMCLIP_ATTN_SPLIT=$v timeout 60 python tools/attn_only.py 512 257 16 >> gpurun_out/split_ab.log 2>&1
done
echo "VLMCLIP_ATTN_SPLIT=3 phase timers (cycles per tile, CTA 0)" >> gpurun_out/split_ab.log
VLMCLIP_ATTN_SPLIT=3 VLMCLIP_ATTN_DEBUG=1 timeout 60 python tools/attn_only.py 512 257 16 2>&1 | grep "attn-pp dbg" | tail -6 >> gpurun_out/split_ab.log
cat gpurun_out/split_ab.log
