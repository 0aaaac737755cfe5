#!/bin/bash
# Round-2 final one-GPU evidence: tests, bench line, reference line, launch list of the bench command, `--set full` captures of
# the dominant kernel (passes A / B / C at res 256), of the volume kernels and of the serial-tail kernels, training numbers.
o=gpurun_out; mkdir -p $o; p=r02f
timeout 600 python bench.py > $o/${p}_bench.txt 2> $o/${p}_bench.err; tail -n 1 $o/${p}_bench.txt > $o/${p}_bench_line.json; cut -c1-300 $o/${p}_bench_line.json
timeout 900 python -m pytest tests -m gpu -q > $o/${p}_pytest_gpu.txt 2>&1; tail -n 3 $o/${p}_pytest_gpu.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $o/${p}_bench_reference.txt 2>&1; tail -n 1 $o/${p}_bench_reference.txt | cut -c1-200
# full captures: the five pass-kernel launches of one batch at res 256 (A, B global, C global, B local, C local)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pointnet_pass_kernel -s 10 -c 5 -f -o $o/${p}_pass_full \
    python bench.py --steps 1 --warmup 1 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/${p}_pass_full.log 2>&1; tail -n 2 $o/${p}_pass_full.log
ncu -i $o/${p}_pass_full.ncu-rep --page raw --csv > $o/${p}_pass_full_raw.csv 2>/dev/null
# launch list of the bench command (per-launch times are cold-cache and serialised: only the shares are used)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv \
    --log-file $o/${p}_launches.csv python bench.py --steps 1 --warmup 1 --cpu_sample 0 --skip_sharded > $o/${p}_launches.log 2>&1
python tools/summarize_launches.py $o/${p}_launches.csv > $o/${p}_launches_summary.txt 2>&1; head -n 20 $o/${p}_launches_summary.txt
# serial tail of one batch: kNN, sub-sample, operand packing, FC layers
timeout 900 ncu --set full --clock-control none -k regex:'subsample_|knn_patch|fc_tc_kernel|pack_a' -s 16 -c 18 -f -o $o/${p}_tail_full \
    python bench.py --steps 1 --warmup 0 --cpu_sample 0 --skip_mesh_stage --skip_sharded > $o/${p}_tail_full.log 2>&1; tail -n 2 $o/${p}_tail_full.log
ncu -i $o/${p}_tail_full.ncu-rep --page raw --csv > $o/${p}_tail_full_raw.csv 2>/dev/null
# volume stage: sign propagation + marching cubes kernels
timeout 900 ncu --set full --clock-control none -k regex:'propagate_kernel|mc_|init_sign|finalize|scatter_kernel' -c 12 -f -o $o/${p}_volume_full \
    python tools/prof_vol.py 256 1 > $o/${p}_volume_full.log 2>&1; tail -n 2 $o/${p}_volume_full.log
ncu -i $o/${p}_volume_full.ncu-rep --page raw --csv > $o/${p}_volume_full_raw.csv 2>/dev/null
P2S_VOL_STATS=1 python tools/prof_vol.py 256 2 > $o/${p}_vol256_stats.txt 2>&1; tail -n 2 $o/${p}_vol256_stats.txt
P2S_VOL_STATS=1 python tools/prof_vol.py 512 2 > $o/${p}_vol512_stats.txt 2>&1; tail -n 2 $o/${p}_vol512_stats.txt
timeout 300 python tools/train_bench.py --batch 1024 --steps 5 --warmup 2 > $o/${p}_train_b1024.txt 2>&1; tail -n 1 $o/${p}_train_b1024.txt | cut -c1-300
timeout 300 python tools/train_bench.py --batch 128 --steps 10 --warmup 3 > $o/${p}_train_b128.txt 2>&1; tail -n 1 $o/${p}_train_b128.txt | cut -c1-300
rm -f $o/${p}_*.ncu-rep
echo done
