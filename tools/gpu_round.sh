#!/usr/bin/env bash
# GPU-box side of the measurement loop, one sub-command per gpurun call (a call may hold at most
# one profiler run, and only after the same command exited 0 without it).  Everything is written
# under gpurun_out/, which gpurun merges back.  Usage, from the build container:
#
#   gpurun --timeout 300 -- 'bash tools/gpu_round.sh check'             # GPU tests + bench line
#   gpurun --timeout 300 -- 'bash tools/gpu_round.sh launches'          # bench, then its ncu launch list
#   gpurun --timeout 300 -- 'bash tools/gpu_round.sh full k1_disp2depth_flat'   # bench, then --set full of one kernel
#   gpurun --timeout 200 -- 'bash tools/gpu_round.sh micro'             # K1 / K1b / K2 micro-benchmarks, pipeline timeline
#
# Back in the container:  python tools/ncu_summary.py gpurun_out/launches.csv
#                         ncu -i gpurun_out/full_<kernel>.ncu-rep --page raw --csv | grep -E 'dram__bytes|gpu__time_duration|sm__warps_active|launch__registers'
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
case "${1:-check}" in
  check)
    timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; tail -15 gpurun_out/gpu_tests.log
    timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err
    python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('frames/s %.0f  ms/step %.4f  e2e %.0f  K1 frac %.3f  launches %d' % (
    d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches']))
print('modes:', {k: (round(v['value']), round(v['ms_per_step'], 4), round(v['host_issue_ms_per_step'], 4)) for k, v in d['modes'].items() if isinstance(v, dict)})
print('parity:', d.get('parity'))
print('kernels:', {m: {k: round(v['ms'], 4) for k, v in t.items()} for m, t in d['kernels'].items()})
print('K4 phases (slowest sequence):', d.get('k4_phase_cycles_slowest_sequence'))
print('assignments:', d.get('assignments'))
PY
    ;;
  launches)
    timeout 120 $BENCH > gpurun_out/plain.log 2>&1 || { echo 'plain run failed'; tail -5 gpurun_out/plain.log; exit 1; }
    timeout 250 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
        --clock-control none -k 'regex:^k[0-9]' -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu.log 2>&1
    python tools/ncu_summary.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt
    ;;
  full)
    K="${2:?kernel name regex}"
    timeout 120 $BENCH > gpurun_out/plain.log 2>&1 || { echo 'plain run failed'; tail -5 gpurun_out/plain.log; exit 1; }
    timeout 250 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 6 -c 3 \
        -f -o "gpurun_out/full_$K" $BENCH > gpurun_out/ncu.log 2>&1
    ls -la "gpurun_out/full_$K.ncu-rep"
    ;;
  micro)
    timeout 60 python tools/bench_k1.py > gpurun_out/k1_micro.json 2> gpurun_out/k1_micro.err; tail -2 gpurun_out/k1_micro.json
    timeout 60 python tools/bench_k1b.py > gpurun_out/k1b_micro.json 2> gpurun_out/k1b_micro.err; cat gpurun_out/k1b_micro.json
    timeout 60 python tools/pipeline_timeline.py 12 5 map > gpurun_out/pipeline_timeline_map.txt 2>&1; tail -8 gpurun_out/pipeline_timeline_map.txt
    timeout 60 python tools/pipeline_timeline.py 12 5 raw > gpurun_out/pipeline_timeline_raw.txt 2>&1; tail -8 gpurun_out/pipeline_timeline_raw.txt
    ;;
  scale)
    # gpurun --gpus N -- 'bash tools/gpu_round.sh scale N [extra bench flags]'
    N="${2:?number of GPUs}"; shift 2
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus "$N" --no-cpu-baseline "$@" > "gpurun_out/bench_n$N.json" 2> "gpurun_out/bench_n$N.err"
    tail -3 "gpurun_out/bench_n$N.err"
    python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f'gpurun_out/bench_n{n}.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'frames/s %.0f  ms/step %.4f  e2e %.0f (%.3f ms/step, copy-only %.3f)' % (
    d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_copy_only_ms_per_step']))
print('modes:', {k: (round(v['value']), round(v['ms_per_step'], 4), round(v['host_issue_ms_per_step'], 4)) for k, v in d['modes'].items() if isinstance(v, dict)})
print('per rank e2e:', [(r['rank'], round(r['ms_per_step'], 3), round(r['h2d_gbs'], 1), round(r['h2d_copy_only_gbs'], 1), r['numa']) for r in d['e2e']['per_rank']])
print('parity:', d['parity']['mismatches'], 'clocks:', d['clocks'])
PY
    ;;
  *) echo "unknown sub-command $1"; exit 2 ;;
esac
