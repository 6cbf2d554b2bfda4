"""Warp-stall samples of one kernel per SOURCE line, from an `ncu --set full --import-source on`
report: the SASS page of the report (per-instruction samples) joined with the line table of the
cubin (nvdisasm -g), because this ncu version prints no metrics on its CUDA source page.

    python tools/ncu_hot_lines.py <report.ncu-rep> <kernel name substring> <translation unit, e.g. k4_track2> [top=40]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kname, tu = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
so = os.path.join(ROOT, 'stereotracking_b200', 'libstereotrack_b200.so')
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(tu + '.') and f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cur, func, table = (None, None), None, collections.defaultdict(list)
for l in dis.split('\n'):
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\.text\.(\S+):', l)
    if m:
        func = m.group(1)
        continue
    if func and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        table[func].append(cur)
fn = [f for f in table if kname in f][0]
lines = table[fn]
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True,
                      text=True).stdout
rows = list(csv.reader(sass.split('\n')))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
col = {k: i for i, k in enumerate(rows[h])}
ins = []
for r in rows[h + 1:]:
    if r and r[0].startswith('0x'):
        ins.append(r)
    elif ins and r and r[0] == 'Kernel Name':
        break                                  # first captured launch only
base = int(ins[0][0], 16)
stalls = [k for k in col if k.startswith('stall_') and 'Not Issued' not in k]
agg, kinds, execd = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
for r in ins:
    idx = (int(r[0], 16) - base) // 16
    if idx >= len(lines):
        continue
    key = lines[idx]
    agg[key] += int(r[col['# Samples']] or 0)
    execd[key] += int(r[col['Instructions Executed']] or 0)
    for k in stalls:
        kinds[key][k[6:]] += int(r[col[k]] or 0)
tot = sum(agg.values())
allk = collections.Counter()
for v in kinds.values():
    allk.update(v)
print(f'{fn[:100]}\n{len(lines)} SASS instructions ({len(lines) * 16} bytes), {tot} stall samples')
print('by reason: ' + ', '.join(f'{k}={v} ({100 * v / max(sum(allk.values()), 1):.0f}%)' for k, v in allk.most_common(8)))
for (fi, li), c in agg.most_common(top):
    t3 = ', '.join(f'{k}={v}' for k, v in kinds[(fi, li)].most_common(3) if v)
    print(f'{fi}:{li:5d}  samples {c:5d} ({100 * c / max(tot, 1):4.1f}%)  warp-instr {execd[(fi, li)]:8d}  {t3}')
