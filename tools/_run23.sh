set -u
mkdir -p gpurun_out
python tools/bench_k4.py 40 C3 64 > gpurun_out/k4_micro_c3.json 2> gpurun_out/k4_micro_c3.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/k4_micro_c3.json').read())
print('C3', d['rep0'], d.get('stage1_fast_path_refusals_by_code'))
for r in d['slowest_sequence_per_frame_top5']: print(r)
PY
python tools/bench_k4.py 40 C4 64 > gpurun_out/k4_micro_c4.json 2> gpurun_out/k4_micro_c4.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/k4_micro_c4.json').read())
print('C4', d['rep0'], d.get('stage1_fast_path_refusals_by_code'))
for r in d['slowest_sequence_per_frame_top5']: print(r)
PY
tail -3 gpurun_out/k4_micro_c4.err
