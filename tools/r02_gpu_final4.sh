#!/bin/bash
# final check of the round without the reference arm and the kernel bench (both unchanged since r02_gpu_final3.sh), plus the
# full fine-tune step through bench.py
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_final_build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_pytest.log
tail -4 gpurun_out/r02_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_final_smoke.log
tail -2 gpurun_out/r02_final_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/r02_final_bench.json
timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 > gpurun_out/r02_cfg5_final.json 2> gpurun_out/r02_cfg5_final.err; echo "cfg5 rc=$?"
cut -c1-500 gpurun_out/r02_cfg5_final.json
