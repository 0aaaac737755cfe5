#!/bin/bash
# Round evidence on one B200 (run through gpurun): tests, bench lines, ncu captures.  Everything lands in gpurun_out/.
# Numbers printed by runs under ncu are never bench values; only per-kernel times / bytes are used from them.
tag=${1:-r01}
mkdir -p gpurun_out
o=gpurun_out
timeout 700 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu.txt 2>&1; tail -n 3 $o/${tag}_pytest_gpu.txt
timeout 500 python bench.py --steps 5 --warmup 3 > $o/${tag}_bench.txt 2>&1; tail -n 1 $o/${tag}_bench.txt > $o/${tag}_bench_line.json; cut -c1-300 $o/${tag}_bench_line.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 --cpu_sample 128 > $o/${tag}_bench_reference.txt 2>&1; tail -n 1 $o/${tag}_bench_reference.txt | cut -c1-300
timeout 200 python bench.py --grid_res 128 --steps 3 --warmup 3 --cpu_sample 0 > $o/${tag}_bench_res128.txt 2>&1
timeout 200 python bench.py --model max --steps 3 --warmup 3 --cpu_sample 0 > $o/${tag}_bench_max256.txt 2>&1
timeout 300 python bench.py --grid_res 512 --steps 2 --warmup 3 --cpu_sample 0 > $o/${tag}_bench_res512.txt 2>&1
for f in res128 max256 res512; do tail -n 1 $o/${tag}_bench_$f.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', round(d['value']), 'q/s', round(d['ms_per_step'],1), 'ms', d['config']['queries_per_shape'], 'Q', d['mesh_stage'])" ; done
timeout 300 python tools/train_bench.py --batch 1024 --steps 5 --warmup 2 --cpu_sample 32 > $o/${tag}_train_b1024.txt 2>&1; tail -n 1 $o/${tag}_train_b1024.txt | cut -c1-400
timeout 200 python tools/train_bench.py --batch 128 --steps 5 --warmup 2 > $o/${tag}_train_b128.txt 2>&1; tail -n 1 $o/${tag}_train_b128.txt | cut -c1-300
# ncu: launch list of one bench step (times only: 1 pass per kernel)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --cpu_sample 0 --skip_mesh_stage > $o/${tag}_launches.log 2>&1
python tools/summarize_launches.py $o/${tag}_launches.csv > $o/${tag}_launches_summary.txt 2>&1; head -n 14 $o/${tag}_launches_summary.txt
# ncu --set full on the pass kernel (5 consecutive launches = one batch), source-level
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pointnet_pass -s 10 -c 5 -o $o/${tag}_pass_full \
    python bench.py --steps 1 --warmup 3 --cpu_sample 0 --grid_res 128 > $o/${tag}_pass_full.log 2>&1
ncu -i $o/${tag}_pass_full.ncu-rep --page raw --csv > $o/${tag}_pass_full_raw.csv 2>/dev/null
# ncu: DRAM bytes + time of the byte kernels (assembly, volume stage, marching cubes) on a res-128 step
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'knn_patch|subsample|gather|scatter|box_|vote|finalize|init_sign|mc_|dilate|occupancy|query_points|sdf_from' -c 1200 --csv \
    --log-file $o/${tag}_byte_kernels.csv python bench.py --steps 1 --warmup 1 --cpu_sample 0 --grid_res 128 > $o/${tag}_byte_kernels.log 2>&1
python tools/summarize_launches.py $o/${tag}_byte_kernels.csv > $o/${tag}_byte_kernels_summary.txt 2>&1; head -n 24 $o/${tag}_byte_kernels_summary.txt
