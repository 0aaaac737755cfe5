#!/bin/bash
o=gpurun_out; mkdir -p $o
P2S_TC_WAITSTATS=1 timeout 300 python bench.py --grid_res 128 --steps 1 --warmup 1 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/r02_waitstats.txt 2>&1
grep -A6 "p2s waitstats" $o/r02_waitstats.txt | head -80
echo done
