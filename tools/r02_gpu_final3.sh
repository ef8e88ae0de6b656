#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_final_build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_pytest.log
tail -4 gpurun_out/r02_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_final_smoke.log
tail -2 gpurun_out/r02_final_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"
cat gpurun_out/r02_final_bench.json
timeout 300 python tools/kernel_bench.py > gpurun_out/r02_kernel_bench.log 2>&1; echo "kernel_bench rc=$?"; tail -30 gpurun_out/r02_kernel_bench.log
