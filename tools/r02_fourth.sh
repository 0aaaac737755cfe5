#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "sign_propagation or all_zero or mesh_chamfer or chamfer_engine" > gpurun_out/r02d_pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02d_pytest.txt; tail -3 gpurun_out/r02d_pytest.txt
P2S_VOL_STATS=1 python tools/prof_vol.py 256 2 2>&1 | tail -4
P2S_VOL_STATS=1 python tools/prof_vol.py 512 2 2>&1 | tail -4
timeout 300 python bench.py --steps 4 --warmup 3 --cpu_sample 0 --skip_mesh_stage --skip_sharded > gpurun_out/r02d_bench.txt 2>&1; grep -o '"value": [0-9.]*' gpurun_out/r02d_bench.txt | head -1
echo done
