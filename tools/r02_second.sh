#!/bin/bash
# round 2, second GPU call: full GPU suite (ball query, 1200-point patches, MC decider, full_eval shim, e2e res 128), bench with the
# sharded sections, sign-propagation statistics at 256^3 / 512^3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02b_pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02b_pytest.txt
tail -5 gpurun_out/r02b_pytest.txt
P2S_VOL_STATS=1 timeout 600 python bench.py > gpurun_out/r02b_bench.txt 2> gpurun_out/r02b_bench.err; tail -c 2500 gpurun_out/r02b_bench.txt; grep "sign propagation" gpurun_out/r02b_bench.err | tail -3
P2S_VOL_STATS=1 timeout 400 python bench.py --grid_res 512 --steps 1 --warmup 3 --cpu_sample 0 --skip_sharded > gpurun_out/r02b_bench512.txt 2> gpurun_out/r02b_bench512.err; grep "sign propagation" gpurun_out/r02b_bench512.err | tail -2
echo done
