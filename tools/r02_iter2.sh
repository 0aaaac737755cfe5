#!/bin/bash
# round 2 iteration call (sub-sampler work): relevant GPU tests, stage timing, short bench, launch list, full capture of the sub-sampler
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -x -k "subsample or reconstruct or headline or plumbing or dropin or slab" > $o/iter_pytest.txt 2>&1; echo "pytest rc $?" >> $o/iter_pytest.txt; tail -n 4 $o/iter_pytest.txt
P2S_STAGE_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 3 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/iter_stage.txt 2>&1; grep -A14 "stage timing" $o/iter_stage.txt | head -20
timeout 300 python bench.py --steps 5 --warmup 3 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/iter_bench.txt 2>&1; grep -o '"value": [0-9.]*' $o/iter_bench.txt | head -2
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $o/iter_launches.csv python bench.py --steps 1 --warmup 0 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/iter_launches.log 2>&1
python tools/summarize_launches.py $o/iter_launches.csv 2>&1 | head -9
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'subsample_cells' -s 1 -c 1 -f -o $o/iter_sub_full \
    python bench.py --steps 1 --warmup 0 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/iter_sub_full.log 2>&1; tail -n 1 $o/iter_sub_full.log
ncu -i $o/iter_sub_full.ncu-rep --page raw --csv > $o/iter_sub_full_raw.csv 2>/dev/null
ncu -i $o/iter_sub_full.ncu-rep --page source --csv > $o/iter_sub_full_src.csv 2>/dev/null
rm -f $o/iter_sub_full.ncu-rep
echo done
