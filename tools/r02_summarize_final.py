"""Collect the round-2 final evidence files from gpurun_out/ into profiles/ and write profiles/r02f_summary.md."""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
keep = ['r02f_pytest_gpu.txt', 'r02f_bench_line.json', 'r02f_bench_reference.txt', 'r02f_launches.csv', 'r02f_launches_summary.txt',
        'r02f_pass_full_raw.csv', 'r02f_tail_full_raw.csv', 'r02f_volume_full_raw.csv', 'r02f_vol256_stats.txt', 'r02f_vol512_stats.txt',
        'r02f_train_b1024.txt', 'r02f_train_b128.txt']
for f in keep:
    p = os.path.join(src, f)
    if os.path.exists(p):
        shutil.copy(p, os.path.join(dst, f))
    else:
        print('missing', f)


def raw(fn, wanted):
    p = os.path.join(src, fn)
    if not os.path.exists(p):
        return []
    rows = list(csv.reader(open(p)))
    if len(rows) < 3:
        return []
    hdr, data = rows[0], rows[2:]
    out = []
    for r in data:
        out.append({'name': r[hdr.index('Kernel Name')].split('(')[0][-44:], **{w: r[hdr.index(w)] for w in wanted if w in hdr}})
    return out


lines = ['# Round 2, final evidence call (1 x B200, `tools/r02_final.sh`)', '']
try:
    b = json.loads(open(os.path.join(src, 'r02f_bench_line.json')).read())
    rf = b['roofline']
    lines += ['| item | result |', '|---|---|',
              '| `python bench.py` value | %.1f k queries/s (%.1f ms per %d-query shape), e2e %.1f k; clocks %s/%s MHz %s |' % (
                  b['value'] / 1e3, b['ms_per_step'], b['config']['queries_per_shape'], b['e2e']['value'] / 1e3, b['clocks']['sm_mhz'], b['clocks']['sm_max_mhz'], b['clocks']['reasons']),
              '| pass kernel (bench CUDA events) | %.3f ms per launch = %.0f TFLOP/s algorithmic = %.3f of the measured sustained bf16 peak, %.1f %% of the step |' % (
                  rf['ms_per_launch'], rf['achieved'], rf['frac'], 100 * rf['share_of_step']),
              '| mesh stage | sign propagation %.2f ms (%d iterations), marching cubes %.2f ms, %d faces |' % (
                  b['mesh_stage']['sign_propagation_ms'], b['mesh_stage']['sign_propagation_iterations'], b['mesh_stage']['marching_cubes_ms'], b['mesh_stage']['faces']),
              '| sharded job (1 GPU) | %s |' % json.dumps({k: v for k, v in b['sharded_job'].items() if k != 'workload'}),
              '| tile-sharded shape (1 GPU) | %s |' % json.dumps({k: v for k, v in b['tile_sharded_one_shape'].items() if k != 'workload'}),
              '| cpu_baseline | %s |' % json.dumps(b['cpu_baseline']), '']
except Exception as e:  # noqa: BLE001
    lines.append('bench line missing: %r' % (e,))
for fn, title in (('r02f_pass_full_raw.csv', 'pass kernel launches of one batch (A, B global, C global, B local, C local), `ncu --set full`'),
                  ('r02f_tail_full_raw.csv', 'serial tail of one batch, `ncu --set full`'),
                  ('r02f_volume_full_raw.csv', 'volume stage at 256^3, `ncu --set full`')):
    w = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
         'sm__inst_executed.avg.per_cycle_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum']
    rows = raw(fn, w)
    if rows:
        lines += ['## ' + title, '', '| kernel | time | tensor pipe active % | tensor operand fetch active % | IPC | l1tex % | DRAM read | DRAM write |', '|---|---|---|---|---|---|---|---|']
        for r in rows:
            lines.append('| `%s` | %s | %s | %s | %s | %s | %s | %s |' % (r['name'], *[r.get(k, '') for k in w]))
        lines.append('')
for fn in ('r02f_launches_summary.txt', 'r02f_vol256_stats.txt', 'r02f_vol512_stats.txt', 'r02f_pytest_gpu.txt', 'r02f_train_b1024.txt', 'r02f_train_b128.txt'):
    p = os.path.join(src, fn)
    if os.path.exists(p):
        lines += ['## ' + fn, '', '```'] + [l.rstrip()[:400] for l in open(p).read().splitlines()[-16:]] + ['```', '']
open(os.path.join(dst, 'r02f_summary.md'), 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines)[:6000])
