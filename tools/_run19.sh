timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -4
ST_K2_PERSIST=1 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "k2" 2>&1 | tail -2
for pz in 0 1; do
  ST_K2_PERSIST=$pz timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k2_score' --csv --log-file gpurun_out/k2_score_persist$pz.csv python tools/bench_k2.py > gpurun_out/ncu_k2_p$pz.log 2>&1
  python - $pz <<'PY'
import csv, sys, collections
pz = sys.argv[1]
rows = [r for r in csv.reader(open(f'gpurun_out/k2_score_persist{pz}.csv')) if len(r) > 5]
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]; gi, mi, vi, ui, ii = hdr.index('Grid Size'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('ID')
per = collections.defaultdict(list)
for r in rows[h + 1:]:
    if r[mi] == 'gpu__time_duration.sum':
        per[int(r[ii]) // 40].append(float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3}[r[ui]])
for g, v in sorted(per.items()):
    v.sort()
    print(f'k2_score persist={pz} launches {g*40}..: n={len(v)} median {v[len(v) // 2]:.1f} us')
PY
done
timeout 400 python bench.py --workload C3 --seq 64 --steps 20 --warmup 4 --no-cpu-baseline --e2e-reps 1 > gpurun_out/bench_c3_s64.json 2> gpurun_out/bench_c3_s64.err; tail -2 gpurun_out/bench_c3_s64.err
timeout 400 python bench.py --workload C4 --seq 64 --steps 20 --warmup 4 --no-cpu-baseline --e2e-reps 1 > gpurun_out/bench_c4_s64.json 2> gpurun_out/bench_c4_s64.err; tail -2 gpurun_out/bench_c4_s64.err
timeout 300 python bench.py --no-cpu-baseline --e2e-reps 1 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
for f in ['bench_c3_s64', 'bench_c4_s64', 'bench']:
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value %.0f (%.4f ms/step)' % (d['value'], d['ms_per_step']), {k: (round(v['value']), round(v['ms_per_step'], 4)) for k, v in d['modes'].items() if isinstance(v, dict)},
              'parity', d['parity']['mismatches'], 'status', d['status_bits'], 'k4 ms', round(d['kernels']['raw']['k4_track_step']['ms'], 3), d['assignments'], d['k4_phase_cycles_slowest_sequence'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
