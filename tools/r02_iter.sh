#!/bin/bash
# round 2 iteration call: GPU tests, stage timing, volume-stage statistics, short bench, short launch list
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -x > $o/iter_pytest.txt 2>&1; echo "pytest rc $?" >> $o/iter_pytest.txt; tail -n 6 $o/iter_pytest.txt
P2S_VOL_STATS=1 python tools/prof_vol.py 256 2 2>&1 | tail -4 | tee $o/iter_vol256.txt
P2S_VOL_STATS=1 python tools/prof_vol.py 512 2 2>&1 | tail -4 | tee $o/iter_vol512.txt
P2S_STAGE_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 3 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/iter_stage.txt 2>&1; grep -A14 "stage timing" $o/iter_stage.txt | head -20
timeout 300 python bench.py --steps 5 --warmup 3 --cpu_sample 0 --skip_sharded > $o/iter_bench.txt 2>&1; grep -o '"value": [0-9.]*' $o/iter_bench.txt | head -2; grep -o '"sign_propagation_ms": [0-9.]*' $o/iter_bench.txt
# per-kernel times of one step (cold-cache, serialised: shares only)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $o/iter_launches.csv python bench.py --steps 1 --warmup 0 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/iter_launches.log 2>&1
python tools/summarize_launches.py $o/iter_launches.csv 2>&1 | head -14
echo done
