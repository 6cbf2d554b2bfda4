set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -k "k2 or decode or nms or pipeline or bench" 2>&1 | tail -4
med() {  # csv label
python - "$1" "$2" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]; mi, vi, ui = hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
v = [float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3}[r[ui]] for r in rows[h + 1:] if r[mi] == 'gpu__time_duration.sum']
for a in range(0, len(v), 40):
    w = sorted(v[a:a + 40]); print(sys.argv[2], f'launches {a}..: n={len(w)} median {w[len(w)//2]:.1f} us')
PY
}
python tools/bench_k2.py > gpurun_out/k2_flat_calls.txt 2>&1; cat gpurun_out/k2_flat_calls.txt | cut -c1-250
for v in "ST_K2_V4=1" "ST_K2_ITEMS=2" "ST_K2_ITEMS=4" "ST_K2_ITEMS=8"; do
  n=$(echo $v | tr '=' '_')
  env $v timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k 'regex:k2_score' --csv \
      --log-file gpurun_out/k2_score_$n.csv python tools/bench_k2.py > gpurun_out/ncu_k2_$n.log 2>&1
  med gpurun_out/k2_score_$n.csv "$v"
done
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k2_score' -s 8 -c 1 -f -o gpurun_out/full_k2_score_flat python tools/bench_k2.py big > gpurun_out/ncu_full_k2.log 2>&1
ncu -i gpurun_out/full_k2_score_flat.ncu-rep --page raw --csv > gpurun_out/full_k2_score_flat_metrics.csv 2>/dev/null
python tools/ncu_hot_lines.py gpurun_out/full_k2_score_flat.ncu-rep k2_score_flat k2_decode_nms 25 > gpurun_out/k2_hot_lines.txt 2>&1; head -40 gpurun_out/k2_hot_lines.txt
