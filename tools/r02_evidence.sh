#!/bin/bash
# Round-2 one-GPU evidence: tests, bench line, launch list of the bench command, `--set full` captures of the dominant kernel
# (passes A / B / C at res 256) and of the volume kernels, launch list of one training step.
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q > $o/r02_pytest_gpu.txt 2>&1; tail -n 3 $o/r02_pytest_gpu.txt
timeout 600 python bench.py > $o/r02_bench.txt 2> $o/r02_bench.err; tail -n 1 $o/r02_bench.txt > $o/r02_bench_line.json; cut -c1-300 $o/r02_bench_line.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/r02_bench_reference.txt 2>&1; tail -n 1 $o/r02_bench_reference.txt | cut -c1-300
# launch list of the bench command (per-launch times are cold-cache and serialised: only the shares are used)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
    --log-file $o/r02_launches.csv python bench.py --steps 2 --warmup 1 --cpu_sample 0 --skip_sharded > $o/r02_launches.log 2>&1
python tools/summarize_launches.py $o/r02_launches.csv > $o/r02_launches_summary.txt 2>&1; head -n 24 $o/r02_launches_summary.txt
# full captures: the five pass-kernel launches of one batch at res 256 (A, B global, C global, B local, C local)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pointnet_pass_kernel -s 10 -c 5 -f -o $o/r02_pass_full \
    python bench.py --steps 1 --warmup 1 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/r02_pass_full.log 2>&1; tail -n 2 $o/r02_pass_full.log
ncu -i $o/r02_pass_full.ncu-rep --page raw --csv > $o/r02_pass_full_raw.csv 2>/dev/null
# volume stage: sign propagation + marching cubes kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'propagate_kernel|mc_|init_sign|finalize|scatter_kernel' -c 12 -f -o $o/r02_volume_full \
    python tools/prof_vol.py 256 1 > $o/r02_volume_full.log 2>&1; tail -n 2 $o/r02_volume_full.log
ncu -i $o/r02_volume_full.ncu-rep --page raw --csv > $o/r02_volume_full_raw.csv 2>/dev/null
# one training step (config 4 shape: 128 queries per rank): launch list
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
    --log-file $o/r02_train_launches.csv python tools/train_bench.py --batch 128 --steps 1 --warmup 1 > $o/r02_train_launches.log 2>&1
python tools/summarize_launches.py $o/r02_train_launches.csv > $o/r02_train_launches_summary.txt 2>&1; head -n 16 $o/r02_train_launches_summary.txt
timeout 300 python tools/train_bench.py --batch 1024 --steps 5 --warmup 2 > $o/r02_train_b1024.txt 2>&1; tail -n 2 $o/r02_train_b1024.txt
timeout 300 python tools/train_bench.py --batch 128 --steps 10 --warmup 3 > $o/r02_train_b128.txt 2>&1; tail -n 2 $o/r02_train_b128.txt
echo done
