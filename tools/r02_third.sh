#!/bin/bash
# round 2, third GPU call: schedule variants of the pass kernel (A/B), sign-propagation timing diagnostics
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -k "bit_for_bit or sign_propagation or forward_tc or reconstruct" > gpurun_out/r02c_pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02c_pytest.txt
tail -3 gpurun_out/r02c_pytest.txt
for opt in 0 1 2 3; do
  P2S_TC_OPT=$opt timeout 300 python bench.py --steps 4 --warmup 3 --cpu_sample 0 --skip_mesh_stage --skip_sharded > gpurun_out/r02c_bench_opt$opt.txt 2>&1
  python - <<PY
import json
for l in open('gpurun_out/r02c_bench_opt$opt.txt'):
    if l.startswith('{'):
        d=json.loads(l); print('opt $opt value %.0f ms/step %.2f pass ms/launch %.4f frac %.3f share %.3f' % (d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['roofline']['share_of_step']))
PY
done
P2S_TC_BCMAX=16384 timeout 300 python bench.py --batch 16384 --steps 4 --warmup 3 --cpu_sample 0 --skip_mesh_stage --skip_sharded > gpurun_out/r02c_bench_bc16k.txt 2>&1
grep -o '"value": [0-9.]*' gpurun_out/r02c_bench_bc16k.txt | head -1
P2S_TC_WAITSTATS=1 timeout 300 python bench.py --grid_res 128 --steps 1 --warmup 3 --cpu_sample 0 --skip_mesh_stage --skip_sharded > gpurun_out/r02c_waitstats.txt 2>&1
P2S_VOL_STATS=1 timeout 300 python bench.py --steps 1 --warmup 3 --cpu_sample 0 --skip_sharded > gpurun_out/r02c_vol.txt 2>&1; grep "sign propagation" gpurun_out/r02c_vol.txt | tail -2
P2S_VOL_STATS=1 timeout 300 python bench.py --grid_res 512 --steps 1 --warmup 3 --cpu_sample 0 --skip_sharded > gpurun_out/r02c_vol512.txt 2>&1; grep "sign propagation" gpurun_out/r02c_vol512.txt | tail -1
echo done
