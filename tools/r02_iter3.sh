#!/bin/bash
# pass-kernel iteration: network parity tests, bench, wait statistics
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -q -x -k "forward or tensor or split or headline or model or reconstruct or patch" > $o/iter_pytest.txt 2>&1; echo "pytest rc $?" >> $o/iter_pytest.txt; tail -n 4 $o/iter_pytest.txt
timeout 300 python bench.py --steps 5 --warmup 3 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/iter_bench.txt 2>&1; grep -o '"value": [0-9.]*' $o/iter_bench.txt | head -2; grep -o '"ms_per_launch": [0-9.]*' $o/iter_bench.txt; grep -o '"sm_mhz": [0-9]*' $o/iter_bench.txt
P2S_TC_WAITSTATS=1 timeout 300 python bench.py --grid_res 128 --steps 1 --warmup 0 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/iter_waitstats.txt 2>&1
grep -A6 "p2s waitstats" $o/iter_waitstats.txt | head -36
echo done
