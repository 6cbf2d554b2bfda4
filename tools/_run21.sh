set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -k "k1b or k2 or decode or nms or pipeline or bench" 2>&1 | tail -6
python tools/bench_k2.py > gpurun_out/k2_flat_calls.txt 2>&1; cat gpurun_out/k2_flat_calls.txt | cut -c1-250
for v in "ST_K2_ITEMS=0" "ST_K2_ITEMS=2" "ST_K2_ITEMS=4" "ST_K2_ITEMS=8"; do
  n=$(echo $v | tr '=' '_')
  env $v timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k 'regex:k2_score' --csv \
      --log-file gpurun_out/k2_score_$n.csv python tools/bench_k2.py > gpurun_out/ncu_k2_$n.log 2>&1
done
python tools/k2_cases.py gpurun_out/k2_score_ST_K2_ITEMS_0.csv gpurun_out/k2_score_ST_K2_ITEMS_2.csv gpurun_out/k2_score_ST_K2_ITEMS_4.csv gpurun_out/k2_score_ST_K2_ITEMS_8.csv | tee gpurun_out/k2_score_flat_cases.txt
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k2_score' -s 8 -c 1 -f -o gpurun_out/full_k2_score_flat python tools/bench_k2.py big > gpurun_out/ncu_full_k2.log 2>&1
ncu -i gpurun_out/full_k2_score_flat.ncu-rep --page raw --csv > gpurun_out/full_k2_score_flat_metrics.csv 2>/dev/null
python tools/ncu_hot_lines.py gpurun_out/full_k2_score_flat.ncu-rep k2_score_flat k2_decode_nms 12 > gpurun_out/k2_hot_lines.txt 2>&1; head -16 gpurun_out/k2_hot_lines.txt
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], d.get('parity'), d['e2e']['value'])
PY
