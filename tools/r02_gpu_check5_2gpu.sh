#!/bin/bash
# Round 2, 2-GPU pass: NCCL parity test, then short benches of the three workloads on 2 ranks.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q --tb=short -x -s > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest5.log
tail -30 gpurun_out/r02_pytest5.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench5_cfg2_2gpu.json 2> gpurun_out/r02_bench5_cfg2_2gpu.err; echo "cfg2 rc=$?"
tail -3 gpurun_out/r02_bench5_cfg2_2gpu.err; cat gpurun_out/r02_bench5_cfg2_2gpu.json
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload cfg3 > gpurun_out/r02_bench5_cfg3_2gpu.json 2> gpurun_out/r02_bench5_cfg3_2gpu.err; echo "cfg3 rc=$?"
tail -3 gpurun_out/r02_bench5_cfg3_2gpu.err; cat gpurun_out/r02_bench5_cfg3_2gpu.json
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload cfg5 > gpurun_out/r02_bench5_cfg5_2gpu.json 2> gpurun_out/r02_bench5_cfg5_2gpu.err; echo "cfg5 rc=$?"
tail -3 gpurun_out/r02_bench5_cfg5_2gpu.err; cat gpurun_out/r02_bench5_cfg5_2gpu.json
