set -u
mkdir -p gpurun_out
ST_PIPE_PARTITION=1 python -m pytest tests -m gpu -q -k "pipeline or bench or graph or lockstep" 2>&1 | tail -4
for part in 0 1; do
  ST_PIPE_PARTITION=$part python bench.py --no-cpu-baseline --reps 5 --e2e-reps 1 > gpurun_out/bench_part$part.json 2> gpurun_out/bench_part$part.err; tail -2 gpurun_out/bench_part$part.err
done
ST_PIPE_PARTITION=1 python tools/pipeline_timeline.py 16 5 raw 1 > gpurun_out/pipeline_timeline_raw_graph_part.txt 2>&1; tail -4 gpurun_out/pipeline_timeline_raw_graph_part.txt
python - <<'PY'
import json
for part in (0, 1):
    d = json.loads(open(f'gpurun_out/bench_part{part}.json').read().strip().splitlines()[-1])
    print('partition', part, d['config'].get('sm_partition'), 'value', round(d['value']), d['ms_per_step'], {k: (round(v['value']), round(v['ms_per_step'], 4)) for k, v in d['modes'].items() if isinstance(v, dict)}, 'e2e', round(d['e2e']['value']), 'parity', d['parity']['mismatches'], d['value_stats']['min_ms_per_step'], d['value_stats']['max_ms_per_step'])
PY
