#!/bin/bash
# full ncu captures of the serial-tail kernels (one batch at res 256): kNN, sub-sample, FC tails
o=gpurun_out; mkdir -p $o
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'subsample_reject|knn_patch|fc_tc_kernel' -s 14 -c 14 -f -o $o/r02_tail_full \
    python bench.py --steps 1 --warmup 0 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/r02_tail_full.log 2>&1; tail -n 2 $o/r02_tail_full.log
ncu -i $o/r02_tail_full.ncu-rep --page raw --csv > $o/r02_tail_full_raw.csv 2>/dev/null
ls -la $o/r02_tail_full.ncu-rep
echo done
