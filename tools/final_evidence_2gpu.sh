#!/bin/bash
# Two-GPU part of the round evidence: weak-scaling bench line (incl. the NCCL mesh gather) and the data-parallel train step.
tag=${1:-r01}
mkdir -p gpurun_out
o=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 3 --warmup 3 --cpu_sample 0 > $o/${tag}_bench_2gpu.txt 2>&1; tail -n 1 $o/${tag}_bench_2gpu.txt | cut -c1-400
tail -n 1 $o/${tag}_bench_2gpu.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['mesh_stage'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    tools/train_bench.py --batch 1024 --steps 5 --warmup 2 > $o/${tag}_train_2gpu.txt 2>&1; tail -n 1 $o/${tag}_train_2gpu.txt | cut -c1-400
