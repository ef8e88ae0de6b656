#!/bin/bash
# Round 2, 2-GPU pass: NCCL parity test, then lean benches (checks that multi-rank runs END: graphs released before teardown).
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 240 python -m pytest tests/test_gpu_dp.py -m gpu -q --tb=short -x -s > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest10.log
grep -v "^E  \|Warn" gpurun_out/r02_pytest10.log | tail -12
grep "AssertionError\|dp_eager\|dp_graph\|DP_OK" gpurun_out/r02_pytest10.log | head -8 | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 240 $TR bench.py --gpus 2 --steps 20 --warmup 3 --lean > gpurun_out/r02_bench10_cfg2_2gpu.json 2> gpurun_out/r02_bench10_cfg2_2gpu.err; echo "cfg2 rc=$?"
tail -2 gpurun_out/r02_bench10_cfg2_2gpu.err | cut -c1-300; cat gpurun_out/r02_bench10_cfg2_2gpu.json
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload cfg3 --lean > gpurun_out/r02_bench10_cfg3_2gpu.json 2> gpurun_out/r02_bench10_cfg3_2gpu.err; echo "cfg3 rc=$?"
tail -2 gpurun_out/r02_bench10_cfg3_2gpu.err | cut -c1-300; cat gpurun_out/r02_bench10_cfg3_2gpu.json
