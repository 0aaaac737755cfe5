#!/bin/bash
# round 2, two-GPU call: bench with the sharded sections, config-3 / config-5 style jobs, DDP training step, host loops under torchrun
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02e_bench_2gpu.txt 2> gpurun_out/r02e_bench_2gpu.err; tail -c 1800 gpurun_out/r02e_bench_2gpu.txt
timeout 400 $TR --master-port 29512 bench.py --gpus 2 --workload sharded --model max --grid_res 256 --shapes_per_gpu 1 > gpurun_out/r02e_config3_2gpu.txt 2>&1; tail -c 900 gpurun_out/r02e_config3_2gpu.txt
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --workload sharded --model vanilla --grid_res 512 --shapes_per_gpu 4 > gpurun_out/r02e_config5_2gpu.txt 2>&1; tail -c 900 gpurun_out/r02e_config5_2gpu.txt
timeout 400 $TR --master-port 29514 tools/dist_smoke.py > gpurun_out/r02e_dist_smoke.txt 2>&1; grep "^dist\|Error\|error" gpurun_out/r02e_dist_smoke.txt | tail -6
timeout 400 $TR --master-port 29515 tools/train_bench.py --batch 256 --steps 5 --warmup 2 > gpurun_out/r02e_train_2gpu_b256.txt 2>&1; tail -3 gpurun_out/r02e_train_2gpu_b256.txt
echo done
