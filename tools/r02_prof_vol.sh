#!/bin/bash
mkdir -p gpurun_out
P2S_VOL_STATS=1 python tools/prof_vol.py 256 2 2>&1 | tail -4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -c 1 -f -o gpurun_out/r02_propagate python tools/prof_vol.py 256 1 > gpurun_out/r02_propagate_ncu.log 2>&1; tail -3 gpurun_out/r02_propagate_ncu.log
ls -la gpurun_out/r02_propagate.ncu-rep
echo done
