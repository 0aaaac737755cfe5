#!/bin/bash
# Short re-run of the one-GPU evidence after a kernel change: tests, bench line, launch list of the queries/s step,
# DRAM bytes of the byte kernels.
tag=${1:-r01}
mkdir -p gpurun_out
o=gpurun_out
timeout 700 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu.txt 2>&1; tail -n 3 $o/${tag}_pytest_gpu.txt
timeout 500 python bench.py --steps 5 --warmup 3 > $o/${tag}_bench.txt 2>&1; tail -n 1 $o/${tag}_bench.txt > $o/${tag}_bench_line.json; cut -c1-200 $o/${tag}_bench_line.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --cpu_sample 0 --skip_mesh_stage > $o/${tag}_launches.log 2>&1
python tools/summarize_launches.py $o/${tag}_launches.csv > $o/${tag}_launches_summary.txt 2>&1; head -n 12 $o/${tag}_launches_summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'knn_patch|subsample|gather|scatter|box_|vote|finalize|init_sign|mc_|dilate|occupancy|query_points|sdf_from' -c 1200 --csv \
    --log-file $o/${tag}_byte_kernels.csv python bench.py --steps 1 --warmup 1 --cpu_sample 0 --grid_res 128 > $o/${tag}_byte_kernels.log 2>&1
python tools/summarize_launches.py $o/${tag}_byte_kernels.csv > $o/${tag}_byte_kernels_summary.txt 2>&1; head -n 22 $o/${tag}_byte_kernels_summary.txt
