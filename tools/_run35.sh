set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/bench_k4.py 40 > gpurun_out/k4_micro.json 2> gpurun_out/k4_micro.err; cut -c1-1400 gpurun_out/k4_micro.json
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
for f in ['bench']:
    d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, 'value', round(d['value']), d['ms_per_step'], 'parity', d['parity']['mismatches'], d.get('assignments'), d.get('k4_phase_cycles_slowest_sequence'), d.get('k4_phase_cycles_median'))
PY
