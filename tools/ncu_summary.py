"""Summarise an `ncu --csv` launch list: per-kernel count, average duration, share, DRAM bytes, occupancy."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = None
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, rows = r, rows[i + 1:]
        break
ki, vi, ui, mi, ii = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Metric Name', 'ID'))
per = collections.defaultdict(dict)
for r in rows:
    name = r[ki].split('(')[0].replace('void (anonymous namespace)::', '').replace('<unnamed>::', '').replace('void ', '')
    per[(r[ii], name)][r[mi]] = (float(r[vi].replace(',', '')), r[ui])
conv = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
agg = collections.defaultdict(list)
for (_, name), m in per.items():
    t, u = m['gpu__time_duration.sum']
    t *= {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(u, 1)
    rd, wr = m.get('dram__bytes_read.sum', (0, 'byte')), m.get('dram__bytes_write.sum', (0, 'byte'))
    occ = m.get('sm__warps_active.avg.pct_of_peak_sustained_active', (0, ''))[0]
    agg[name].append((t, rd[0] * conv[rd[1]], wr[0] * conv[wr[1]], occ))
tot = sum(sum(x[0] for x in v) for v in agg.values())
print(f'{"kernel":44s} {"n":>4s} {"avg us":>9s} {"share":>6s} {"dram rd MB":>10s} {"dram wr MB":>10s} {"warps act %":>11s}')
for k, v in sorted(agg.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
    n = len(v)
    print(f'{k[:44]:44s} {n:4d} {sum(x[0] for x in v) / n:9.1f} {sum(x[0] for x in v) / tot * 100:5.1f}% '
          f'{sum(x[1] for x in v) / n / 1e6:10.2f} {sum(x[2] for x in v) / n / 1e6:10.2f} {sum(x[3] for x in v) / n:11.1f}')
