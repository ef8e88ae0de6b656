#!/bin/bash
# First GPU contact of the experimental single-launch key-range attention (csrc/attention_kr.cu, VLMCLIP_ATTN_SPLIT=4).
# Every step runs under its own short timeout: a hung mbarrier wait must not take the box with it.
#   gpurun --timeout 240 -- 'bash tools/gpu_kr_bringup.sh'
mkdir -p gpurun_out
log=gpurun_out/kr_bringup.log
: > $log
echo "== tiny shape under compute-sanitizer (memcheck)" >> $log
VLMCLIP_ATTN_SPLIT=4 timeout 90 compute-sanitizer --tool memcheck python tools/attn_only.py 2 257 2 >> $log 2>&1; echo "rc=$?" >> $log
echo "== parity of variant 4 against the oracle" >> $log
VLMCLIP_RUN_EXPERIMENTAL=1 timeout 90 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "variants_subprocess and 4" >> $log 2>&1; echo "rc=$?" >> $log
echo "== A/B at the config-3 shape" >> $log
for v in 4 3; do
  echo "VLMCLIP_ATTN_SPLIT=$v" >> $log
  VLMCLIP_ATTN_SPLIT=$v timeout 60 python tools/attn_only.py 512 257 16 >> $log 2>&1; echo "rc=$?" >> $log
done
tail -40 $log
