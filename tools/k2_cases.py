"""Per-case medians of the K2 score pass from ncu launch lists of tools/bench_k2.py (5 cases of 35 / 35 / 35 / 35 / 15
launches, the first 5 of each are warm-up):  python tools/k2_cases.py list.csv [...]"""
import csv
import sys

CASES = [('C2 x64 background', 0, 35), ('C2 x64 ~300 cand', 35, 70), ('256 x 1080p', 70, 105), ('512 x 1080p', 105, 140),
         ('2048 x 1080p', 140, 155)]
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[h]
    mi, vi, ui, gi, ki = (hdr.index(k) for k in ('Metric Name', 'Metric Value', 'Metric Unit', 'Grid Size', 'Kernel Name'))
    t, b, g, names = [], [], [], []
    for r in rows[h + 1:]:
        v = float(r[vi].replace(',', ''))
        if r[mi] == 'gpu__time_duration.sum':
            t.append(v * {'ns': 1e-3, 'us': 1, 'ms': 1e3}[r[ui]]); g.append(r[gi]); names.append(r[ki].split('<')[0].split('(')[0][-24:])
        if r[mi] == 'dram__bytes_read.sum':
            b.append(v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[r[ui]])
    print(path)
    for name, a, e in CASES:
        w, bb = sorted(t[a + 5:e]), sorted(b[a + 5:e])
        if not w:
            continue
        us, mb = w[len(w) // 2], bb[len(bb) // 2] / 1e6
        print(f'  {name:20s} grid {g[a + 5]:16s} median {us:7.1f} us  dram read {mb:7.1f} MB  -> {mb / us:6.2f} TB/s')
