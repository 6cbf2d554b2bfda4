set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
python tools/bench_k4.py 40 > gpurun_out/k4_micro.json 2> gpurun_out/k4_micro.err; cat gpurun_out/k4_micro.json | cut -c1-1500
python tools/bench_k4.py 40 C3 64 > gpurun_out/k4_micro_c3.json 2> gpurun_out/k4_micro_c3.err; cat gpurun_out/k4_micro_c3.json | cut -c1-2500
for w in C3 C4; do
  python bench.py --workload $w --seq 64 --no-cpu-baseline > gpurun_out/bench_${w,,}_s64.json 2> gpurun_out/bench_${w,,}_s64.err
done
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
for f in ['bench_c3_s64', 'bench_c4_s64', 'bench']:
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', round(d['value']), d['ms_per_step'], 'parity', d['parity']['mismatches'], 'k4 ms', d.get('k4_alone_ms'), d.get('assignments'), d.get('k4_phase_cycles_slowest_sequence'))
    except Exception as e:
        print(f, 'failed', e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k 'regex:k2_score' --csv \
      --log-file gpurun_out/k2_score_shipped.csv python tools/bench_k2.py > gpurun_out/k2_calls_shipped.txt 2>&1
python tools/k2_cases.py gpurun_out/k2_score_shipped.csv | tee gpurun_out/k2_score_shipped_cases.txt
cut -c1-220 gpurun_out/k2_calls_shipped.txt
