"""Summarise an `ncu --csv` launch list: per kernel launches, total/avg time, share, DRAM bytes and GB/s."""
import csv, sys, re, collections

def main(path):
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    per = collections.OrderedDict()
    for r in rd:
        key = (r['ID'],)
        d = per.setdefault(r['ID'], {'name': r['Kernel Name']})
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        m = r['Metric Name']
        if m == 'gpu__time_duration.sum':
            scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6}[unit]
            d['us'] = v * scale
        else:
            scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
            d[m] = v * scale
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in per.values():
        name = re.sub(r'\(.*', '', d['name'])
        name = name.replace('p2s::', '').replace('(anonymous namespace)::', '')
        a = agg[name]
        a[0] += 1; a[1] += d.get('us', 0.0)
        a[2] += d.get('dram__bytes_read.sum', 0.0); a[3] += d.get('dram__bytes_write.sum', 0.0)
    tot = sum(a[1] for a in agg.values())
    print(f'total {tot/1e3:.1f} ms over {sum(a[0] for a in agg.values())} launches')
    print('| kernel | launches | total ms | share | avg us | DRAM rd MB/launch | DRAM wr MB/launch | DRAM GB/s |')
    print('|---|---|---|---|---|---|---|---|')
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        n, us, rd_, wr = a
        if us / tot < 0.001: continue
        gbs = (rd_ + wr) / (us * 1e-6) / 1e9 if us > 0 else 0
        print(f'| `{name[:70]}` | {n} | {us/1e3:.2f} | {100*us/tot:.1f} % | {us/n:.1f} | {rd_/n/1e6:.2f} | {wr/n/1e6:.2f} | {gbs:.0f} |')

if __name__ == '__main__':
    main(sys.argv[1])
