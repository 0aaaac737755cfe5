#!/bin/bash
# round 2, first GPU call: full GPU test suite, bench line, wait statistics of the pass kernel, stage timing, 512^3 mesh stage
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s -k "not (chamfer and 128)" > gpurun_out/r02a_pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02a_pytest.txt
tail -5 gpurun_out/r02a_pytest.txt
timeout 400 python bench.py > gpurun_out/r02a_bench.txt 2> gpurun_out/r02a_bench.err; tail -c 1500 gpurun_out/r02a_bench.txt
P2S_STAGE_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 3 --cpu_sample 0 --skip_mesh_stage > gpurun_out/r02a_stage.txt 2>&1
P2S_TC_WAITSTATS=1 timeout 300 python bench.py --grid_res 128 --steps 1 --warmup 3 --cpu_sample 0 --skip_mesh_stage > gpurun_out/r02a_waitstats.txt 2>&1
timeout 400 python bench.py --grid_res 512 --steps 1 --warmup 3 --cpu_sample 0 > gpurun_out/r02a_bench512.txt 2>&1
echo done
