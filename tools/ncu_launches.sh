#!/bin/bash
# Launch list of one bench step (times + DRAM bytes per launch).  Usage on the GPU box: tools/ncu_launches.sh <tag>
# Numbers printed by a run under ncu are never bench values; only the per-kernel shares / bytes are used.
tag=${1:-launches}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/${tag}.csv python bench.py --steps 1 --warmup 3 --cpu_sample 0 > gpurun_out/${tag}.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}.csv > gpurun_out/${tag}_summary.txt
tail -n 40 gpurun_out/${tag}_summary.txt
